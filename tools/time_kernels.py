#!/usr/bin/env python
"""Correctness + CUDA-event timing of the hot kernels at BASELINE sizes (one line per case)."""
import sys

import torch

sys.path.insert(0, ".")
from scanner_b200 import cabi, kernels  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(5)
L = cabi.lib()
for n, h, w, pitch in [(64, 1080, 1920, 2048), (8, 1080, 1920, 1920), (16, 2160, 3840, 3840), (64, 480, 640, 640)]:
    surf = torch.randint(0, 256, (n, h * 3 // 2, pitch), dtype=torch.uint8, device="cuda", generator=g)
    plan = kernels.ResizePlan(w, h, 224, 224)
    hist, res = kernels.nv12_hist_resize(surf, w, h, 224, 224, plan)
    rgb = kernels.nv12_to_rgb(surf, w, h)
    h2 = kernels.histogram(rgb)
    ref = torch.stack([torch.stack([torch.bincount((rgb[i, :, :, c] >> 4).flatten().int(), minlength=16)
                                    for c in range(3)]) for i in range(n)]).int()
    ok = bool((hist == h2).all()) and bool((res == kernels.resize(rgb, 224, 224)).all()) and bool((h2 == ref).all())
    L.scn_prof_enable(1)
    for _ in range(10):
        kernels.nv12_hist_resize(surf, w, h, 224, 224, plan)
        kernels.histogram(rgb)
        kernels.nv12_to_rgb(surf, w, h)
    torch.cuda.synchronize()
    p = cabi.prof_report()
    L.scn_prof_enable(0)
    ms = {k: v["ms"] / v["launches"] for k, v in p.items()}
    out = {"n": n, "h": h, "w": w, "ok": ok, "ms": {k: round(v, 4) for k, v in ms.items()}}
    for k, v in ms.items():
        if k.startswith("nv12_hist"):
            out["nv12_hist_GBs"] = round(n * h * w * 1.5 / v / 1e6, 1)
        if k.startswith("hist16"):
            out["rgb_hist_GBs"] = round(n * h * w * 3 / v / 1e6, 1)
        if k.startswith("nv12_to_rgb"):
            out["nv12_to_rgb_GBs"] = round(n * h * w * 4.5 / v / 1e6, 1)
    print(out)
