import sys, torch
sys.path.insert(0, ".")
from scanner_b200 import kernels
g = torch.Generator(device="cuda").manual_seed(5)
n,h,w,pitch=3,1080,1920,2048
surf = torch.randint(0, 256, (n, h * 3 // 2, pitch), dtype=torch.uint8, device="cuda", generator=g)
hist,_ = kernels.nv12_hist_resize(surf, w, h, 224, 224, want_resize=False)
rgb = kernels.nv12_to_rgb(surf, w, h)
h2 = kernels.histogram(rgb)
import numpy as np
np.set_printoptions(linewidth=200)
for i in range(n): print((hist-h2)[i].cpu().numpy())
hist3,_ = kernels.nv12_hist_resize(surf, w, h, 224, 224, want_resize=False)
print('rerun equal', bool((hist3==hist).all()))
print(hist[0].sum(1).cpu().numpy(), h2[0].sum(1).cpu().numpy())
