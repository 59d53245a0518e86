#!/usr/bin/env python
"""Per-kernel CUDA-event timing of the Farneback OpticalFlow path on 1080p pairs."""
import sys

import torch

sys.path.insert(0, ".")
from scanner_b200 import cabi, kernels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
h, w = 1080, 1920
g = torch.Generator(device="cuda").manual_seed(3)
a = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
b = torch.roll(a, shifts=(1, 2), dims=(1, 2)).contiguous()
ws = None
for _ in range(2):
    kernels.optical_flow(a, b)
torch.cuda.synchronize()
L = cabi.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    kernels.optical_flow(a, b)
e1.record()
torch.cuda.synchronize()
print("ms per pair (no profiling):", e0.elapsed_time(e1) / 5 / n)
seq = torch.cat([a, a[:1]])
for _ in range(2):
    kernels.optical_flow_sequence(seq)
e0.record()
for _ in range(5):
    kernels.optical_flow_sequence(seq)
e1.record()
torch.cuda.synchronize()
print("ms per pair, consecutive pairs of one clip (each frame expanded once):", e0.elapsed_time(e1) / 5 / n)
L.scn_prof_enable(1)
for _ in range(3):
    kernels.optical_flow(a, b)
torch.cuda.synchronize()
p = cabi.prof_report()
L.scn_prof_enable(0)
tot = sum(v["ms"] for v in p.values())
for k, v in sorted(p.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:28s} {v['ms'] / 3 / n:8.4f} ms/pair  {v['launches'] // 3 // n:3d} launches/pair  {100 * v['ms'] / tot:5.1f} %")
