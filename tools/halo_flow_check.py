#!/usr/bin/env python
"""BASELINE configs[3] check under torchrun (one rank per GPU, nccl): one clip split into contiguous
intervals, one boundary frame exchanged per internal boundary over NVLink, dense optical flow per
rank; the gathered result must equal a single-GPU run over the whole clip.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scanner_b200 import engine as E, halo, kernels  # noqa: E402


def main():
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    h, w, n, gop = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (1080, 1920, 48, 12)))
    rng = np.random.default_rng(7)  # same clip on every rank
    yuv = rng.integers(0, 256, (n, h * w * 3 // 2), dtype=np.uint8)
    data = E.h264_synth(yuv, w, h, gop=gop)
    eng = E.Engine(gpus=[local])
    sid = eng.add_h264(data)
    halo.sharded_optical_flow(eng, sid, local)  # warm-up (decoder creation, NCCL connect)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    a, flows = halo.sharded_optical_flow(eng, sid, local)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    # halo cost on its own
    frames = eng.decode_to_device(sid, range(*halo.interval_of(n, rank, world)), local)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        halo.exchange_halo(frames, 0, 1)
    e1.record()
    torch.cuda.synchronize()
    halo_ms = e0.elapsed_time(e1) / 10
    # gather on rank 0 and compare with the single-GPU result
    sizes = [halo.interval_of(n, r, world) for r in range(world)]
    parts = [torch.empty((b - a_, h, w, 2), dtype=torch.float32, device=flows.device) for a_, b in sizes]
    # all_gather needs equal shapes: pad to the longest interval
    longest = max(b - a_ for a_, b in sizes)
    padded = torch.zeros((longest, h, w, 2), dtype=torch.float32, device=flows.device)
    padded[:flows.shape[0]] = flows
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    ok = None
    if rank == 0:
        full = torch.cat([g[:b - a_] for g, (a_, b) in zip(gathered, sizes)])
        allf = eng.decode_to_device(sid, range(n), local)
        nxt = torch.cat([allf[1:], allf[-1:]])
        ref = torch.cat([kernels.optical_flow(allf[i:i + 8], nxt[i:i + 8]) for i in range(0, n, 8)])
        ok = bool(torch.equal(full, ref))
        print(json.dumps({"config": "configs[3] OpticalFlow stencil [0,1], one clip sharded by interval",
                          "n_gpus": world, "frames": n, "frame": [h, w], "identical_to_single_gpu": ok,
                          "frames_per_s": n / dt, "wall_s": dt, "halo_bytes_per_boundary": h * w * 3,
                          "halo_exchange_ms": halo_ms}), flush=True)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
    return 0 if ok in (None, True) else 1


if __name__ == "__main__":
    sys.exit(main())
