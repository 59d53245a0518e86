#!/usr/bin/env python
"""Measure end-to-end engine throughput (host H.264 -> NVDEC -> Histogram+Resize -> host rows)
for a few instance counts and stream shapes.  Prints one JSON line per configuration."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from scanner_b200 import engine as E  # noqa: E402
from scanner_b200 import protolite  # noqa: E402

STD = protolite.parse_proto(open("scanner_b200/csrc/ops/stdlib_args.proto").read())


def clip(seed, n, h, w, gop, mode):
    rng = np.random.default_rng(seed)
    k = (n + gop - 1) // gop if mode == "skip" else n
    yuv = rng.integers(0, 256, (k, h * w * 3 // 2), dtype=np.uint8)
    return E.h264_synth(yuv, w, h, gop=gop, non_key=mode, frames=n)


def main():
    E.load_stdlib()
    h, w = 1080, 1920
    n_clips, n_frames, gop = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[4]) if len(sys.argv) > 4 else 60, 30
    for mode in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("skip", "pcm")):
        t0 = time.time()
        clips = [clip(i, n_frames, h, w, gop, mode) for i in range(n_clips)]
        gen_s = time.time() - t0
        for inst in [int(x) for x in (sys.argv[3].split(',') if len(sys.argv) > 3 else '1,2,4,8,14'.split(','))]:
            eng = E.Engine(gpus=[0], instances_per_gpu=inst)
            sids = [eng.add_h264(c) for c in clips]
            g = E.Graph()
            src = g.add_source(True)
            hs = g.add_op("Histogram", [(src, "frame")], device=1)
            rz = g.add_op("Resize", [(src, "frame")], device=1)
            g.add_sink((hs, "histogram"))
            g.add_sink((rz, "frame"))
            jobs = []
            for s in sids:
                j = E.Job()
                j.bind_source(src, s)
                j.set_stream_args(rz, protolite.encode(STD["ResizeArgs"], {"width": 224, "height": 224}))
                jobs.append(j)
            eng.run(g, jobs, 30, 60)  # warm-up (decoder creation, pools)
            t0 = time.time()
            eng.run(g, jobs, 30, 60)
            dt = time.time() - t0
            st = eng.stats()
            print(json.dumps({"mode": mode, "instances": inst, "clips": n_clips, "frames": n_clips * n_frames,
                              "fps": n_clips * n_frames / dt, "wall_s": dt, "bytes_per_clip": len(clips[0]),
                              "gen_s": gen_s, "get_frames_ms": st["intervals_ms"].get("get_frames"),
                              "eval_hist_ms": st["intervals_ms"].get("evaluate:Histogram"),
                              "marshal_ms": st["intervals_ms"].get("op_marshal"),
                              "nvdec_us": {k: v for k, v in st["counters"].items() if k.startswith("nvdec_")}}),
                  flush=True)
            eng.close()


if __name__ == "__main__":
    main()
