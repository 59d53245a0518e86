#!/usr/bin/env python
"""Python kernels over raw frames -- runs anywhere (no GPU, no video decode).

    python examples/python_ops_cpu.py

Shows: a function kernel reading its init arguments from config.args, a batched class kernel with
init and per-stream arguments, a stencil kernel, two jobs chained through a named stream, typed
loading, save_mp4.
"""
import os
import sys
import tempfile
from typing import Any, Sequence

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scanner_b200 as sp  # noqa: E402


@sp.register_python_op()
def Brightness(config, frame: sp.FrameType) -> Any:
    return float(frame.mean()) * config.args.get("gain", 1.0)


@sp.register_python_op(batch=8)
class Tint(sp.Kernel):
    def __init__(self, config, channel=0):
        self.channel = channel

    def new_stream(self, amount=0):
        self.amount = amount

    def execute(self, frame: Sequence[sp.FrameType]) -> Sequence[sp.FrameType]:
        out = []
        for f in frame:
            g = f.astype(np.int32)
            g[..., self.channel] += self.amount
            out.append(g.clip(0, 255).astype(np.uint8))
        return out


@sp.register_python_op(stencil=[-1, 0, 1])
def Smooth(config, value: Sequence[Any]) -> Any:
    known = [v for v in value if v is not None]
    return sum(known) / len(known)


def main():
    frames = np.stack([np.full((48, 64, 3), 40 + 4 * i, np.uint8) for i in range(30)])
    with sp.Client(gpus=[], cpu_instances=2) as sc:
        video = sp.NamedVideoStream(sc, "clip", frames=frames)
        col = sc.io.Input([video])
        tinted = sc.ops.Tint(frame=col, channel=2, amount=[60], batch=8)
        tinted_stream = sp.NamedVideoStream(sc, "clip_tinted")
        brightness = sp.NamedStream(sc, "brightness")
        sc.run([sc.io.Output(tinted, [tinted_stream]),
                sc.io.Output(sc.ops.Brightness(frame=tinted, gain=0.5), [brightness])],
               sp.PerfParams.manual(8, 16))   # (every sink of one job gets the same number of rows)
        # a second job consumes the first job's output: every third value, smoothed over its neighbours
        smooth = sp.NamedStream(sc, "brightness_smooth")
        sampled = sc.streams.Stride(sc.io.Input([brightness]), [3])
        sc.run(sc.io.Output(sc.ops.Smooth(value=sampled), [smooth]), sp.PerfParams.manual(2, 4))
        values = list(smooth.load())
        path = tinted_stream.save_mp4(os.path.join(tempfile.mkdtemp(), "tinted"), fps=15)
    print("smoothed brightness of every third frame:", [round(v, 1) for v in values])
    print("wrote", path, os.path.getsize(path), "bytes")
    return values, path


if __name__ == "__main__":
    main()
