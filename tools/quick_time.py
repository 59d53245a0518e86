import json, sys, torch
sys.path.insert(0, ".")
from scanner_b200 import cabi, kernels
g = torch.Generator(device="cuda").manual_seed(5)
n, h, w, pitch = 64, 1080, 1920, 2048
surf = torch.randint(0, 256, (n, h * 3 // 2, pitch), dtype=torch.uint8, device="cuda", generator=g)
plan = kernels.ResizePlan(w, h, 224, 224)
rgb = kernels.nv12_to_rgb(surf, w, h)
ref_h = kernels.histogram(rgb); ref_r = kernels.resize(rgb, 224, 224)
out = {}
for name, wr in (("hist", False), ("hist+resize", True)):
    hist, res = kernels.nv12_hist_resize(surf, w, h, 224, 224, plan, want_resize=wr)
    ok = bool((hist == ref_h).all()) and (not wr or bool((res == ref_r).all()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(12):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kernels.nv12_hist_resize(surf, w, h, 224, 224, plan, want_resize=wr); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); us = ts[len(ts)//2]
    out[name] = {"ok": ok, "us": round(us, 1), "TBs": round(n * h * w * 1.5 / us / 1e6, 3)}
print(json.dumps(out))
