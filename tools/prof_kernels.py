#!/usr/bin/env python
"""Tiny driver for ncu captures: runs each hot kernel a few times on BASELINE-sized inputs.
usage: python tools/prof_kernels.py [hist|fused|fusedhist|resize|blur|nv12|all] [iters]
(fusedhist: NV12 -> Histogram only, fused: NV12 -> Histogram + Resize)"""
import sys

import torch

sys.path.insert(0, ".")
from scanner_b200 import kernels  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator(device="cuda").manual_seed(1)
n, h, w = 64, 1080, 1920
if which in ("hist", "resize", "blur", "all"):
    frames = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
if which in ("fused", "fusedhist", "nv12", "all"):
    surf = torch.randint(0, 256, (n, h * 3 // 2, 2048), dtype=torch.uint8, device="cuda", generator=g)
plan = kernels.ResizePlan(w, h, 224, 224)
for _ in range(iters):
    if which in ("hist", "all"):
        kernels.histogram(frames)
    if which in ("resize", "all"):
        kernels.resize(frames, 224, 224, plan)
    if which in ("blur", "all"):
        kernels.blur(frames[:8], 3)
    if which in ("nv12", "all"):
        kernels.nv12_to_rgb(surf[:16], w, h)
    if which in ("fused", "all"):
        kernels.nv12_hist_resize(surf, w, h, 224, 224, plan)
    if which == "fusedhist":
        kernels.nv12_hist_resize(surf, w, h, 224, 224, plan, want_resize=False)
torch.cuda.synchronize()
