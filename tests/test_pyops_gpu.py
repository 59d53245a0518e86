"""Python kernels downstream of the GPU stages: NVDEC-decoded frames and GPU op outputs reach a
Python kernel in host memory whichever device type it is scheduled as (the reference's
python_kernel.cpp copies its inputs to the CPU the same way)."""
from typing import Any

import numpy as np
import pytest

import oracle
import scanner_b200 as sp
from scanner_b200 import FrameType, register_python_op
from scanner_b200 import engine as E
from scanner_b200 import types as T

pytestmark = pytest.mark.gpu


@register_python_op(name="PyChannelSums")
def channel_sums(config, frame: FrameType) -> Any:
    return [int(v) for v in frame.reshape(-1, 3).sum(axis=0)]


@register_python_op(name="PyChannelSumsOnGpuInstance", device_type=sp.DeviceType.GPU)
def channel_sums_gpu(config, frame: FrameType) -> Any:
    assert config.devices[0][0] == int(sp.DeviceType.GPU)
    return [int(v) for v in frame.reshape(-1, 3).sum(axis=0)]


@register_python_op(name="PyHistTotal")
def hist_total(config, hist: T.Histogram) -> Any:
    return [int(h.sum()) for h in hist]


def test_python_ops_after_nvdec_and_after_a_gpu_op():
    h, w, n, gop = 96, 128, 12, 4
    rng = np.random.default_rng(3)
    yuv = rng.integers(0, 256, (n, h * w * 3 // 2), dtype=np.uint8)
    stream = E.h264_synth(yuv, w, h, gop=gop)
    want = []
    for i in range(n):
        y = yuv[i, :h * w].reshape(h, w)
        u = yuv[i, h * w:h * w + h * w // 4].reshape(h // 2, w // 2)
        v = yuv[i, h * w + h * w // 4:].reshape(h // 2, w // 2)
        chroma = np.empty((h // 2, w), np.uint8)
        chroma[:, 0::2], chroma[:, 1::2] = u, v
        want.append(oracle.nv12_to_rgb(y, chroma))
    sc = sp.Client(gpus=[0], instances_per_gpu=2)
    frames = sc.io.Input([sp.NamedVideoStream(sc, "clip", data=stream)])
    outs = [sp.NamedStream(sc, name) for name in ("cpu_sums", "gpu_sums", "totals")]
    cols = [sc.ops.PyChannelSums(frame=frames),
            sc.ops.PyChannelSumsOnGpuInstance(frame=frames, device=sp.DeviceType.GPU),
            sc.ops.PyHistTotal(hist=sc.ops.Histogram(frame=frames, device=sp.DeviceType.GPU))]
    sc.run([sc.io.Output(c, [o]) for c, o in zip(cols, outs)], sp.PerfParams.manual(4, 8))
    expect = [[int(x) for x in f.reshape(-1, 3).sum(axis=0)] for f in want]
    assert list(outs[0].load()) == expect
    assert list(outs[1].load()) == expect
    assert list(outs[2].load()) == [[h * w] * 3] * n
    sc.stop()
