#!/usr/bin/env python
"""bench.py -- BASELINE.json configs[1]: 1080p decode + Resize(224) + Histogram, frames/s.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm (rank 0)

A "step" is one pass of the hot path over one batch of `--batch` decoded 1080p surfaces (NV12,
NVDEC layout, pitch 2048): surfaces -> RGB -> {Histogram 3x16 int32, Resize 224x224 RGB24}.
  value   frames/s over all ranks with the surfaces already resident in HBM (CUDA events, max
          over ranks); inputs rotate between two batches, each larger than the 126 MB L2.
  e2e     same metric through the public op call with HOST (pinned) surfaces: H2D of the batch,
          the kernels, D2H of histograms + resized frames, all inside the timed region.
  roofline  dominant kernel's algorithmic bytes / its CUDA-event duration (scn_prof_*), against
          MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the oracle (CPU restatement of the reference arithmetic) on a bounded sample of
          the same surfaces on this box's host cores (rank 0, N=1 only).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, PITCH = 1920, 1080, 2048
DW, DH = 224, 224
SURF_ROWS = H * 3 // 2
B_ALG_FUSED = W * H * 3 // 2 + DW * DH * 3 + 192      # 3,261,120 B / frame
B_ALG_HIST_NV12 = W * H * 3 // 2 + 192                  # the histogram kernel alone
METRIC = "frames/sec (1080p H.264 decode+resize+histogram)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # clocks are evidence, not a dependency
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def make_surfaces_np(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 256, (n, SURF_ROWS, PITCH), dtype=np.uint8)
    return s


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle port of the path (image.cu NV12->RGB, test_ops.cpp Histogram + Resize)
    on all host threads.  Each step = `sample` frames of the same workload."""
    if rank != 0:
        return 0
    import numpy as np
    import oracle
    oracle.lib()
    cores = os.cpu_count() or 1
    sample = max(cores, 8)
    surf = make_surfaces_np(min(sample, 16), 1234)

    def work(i):
        s = surf[i % len(surf)]
        oracle.nv12_hist_resize(s[:H], s[H:], DW, DH, W)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(cores) as ex:
        for _ in range(max(args.warmup, 1)):
            list(ex.map(work, range(sample)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(work, range(sample)))
        dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, sample),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} surfaces/step x {args.steps} steps, oracle C port "
                                       "(NV12->RGB + Histogram + Resize 224), one surface per thread"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, batch):
    return {"workload": "configs[1]: 1080p NV12 decoder surfaces -> Resize(224x224)+Histogram (fused), "
                        "frames device-resident", "frame": [H, W], "pitch": PITCH, "resize": [DH, DW],
            "batch_frames_per_step": batch, "l2_policy": "two rotating input batches, each > L2 (126 MB)",
            "decode": "surfaces synthesised directly (uniform random NV12); NVDEC stage not in this number"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames for the cpu_baseline leg (0=auto)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    from scanner_b200 import cabi, kernels

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B = args.batch
    L = cabi.lib()

    # --- device-resident inputs: two batches of B surfaces, seeded per rank (different "clips")
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    batches = [torch.randint(0, 256, (B, SURF_ROWS, PITCH), dtype=torch.uint8, device=dev, generator=g)
               for _ in range(2)]
    plan = kernels.ResizePlan(W, H, DW, DH, dev)

    def step(i):
        return kernels.nv12_hist_resize(batches[i & 1], W, H, DW, DH, plan)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = L.scn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.scn_launch_count() - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # --- roofline leg: same steps with per-kernel events (scn_prof_*), not used for `value`
    L.scn_prof_enable(1)
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    prof = cabi.prof_report()
    L.scn_prof_enable(0)

    # --- e2e leg: pinned host surfaces in, host results out, copies inside the timed region
    host_in = [torch.from_numpy(make_surfaces_np(B, 77 + rank + 10 * j)).pin_memory() for j in range(2)]
    dev_in = torch.empty((B, SURF_ROWS, PITCH), dtype=torch.uint8, device=dev)
    host_hist = torch.empty((B, 3, 16), dtype=torch.int32).pin_memory()
    host_res = torch.empty((B, DH, DW, 3), dtype=torch.uint8).pin_memory()

    def e2e_step(i):
        dev_in.copy_(host_in[i & 1], non_blocking=True)
        hist, res = kernels.nv12_hist_resize(dev_in, W, H, DW, DH, plan)
        host_hist.copy_(hist, non_blocking=True)
        host_res.copy_(res, non_blocking=True)

    for i in range(3):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        e2e_step(i)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)

    # --- max over ranks
    if dist is not None:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = t[0].item(), t[1].item()

    if rank == 0:
        frames = B * args.steps * world
        value = frames / (ms * 1e-3)
        peak, peak_kind = peaks()
        kname = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
        roof = None
        if kname:
            per_launch_s = prof[kname]["ms"] * 1e-3 / prof[kname]["launches"]
            alg = (B_ALG_HIST_NV12 if kname == "nv12_hist_kernel" else B_ALG_FUSED) * B
            ach = alg / per_launch_s / 1e9
            roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "peak_kind": peak_kind + " (burst copy, MEASURED_PEAKS.json)",
                    "alg_bytes_per_launch": alg, "ms_per_launch": per_launch_s * 1e3, "traffic": None,
                    "kernel_share_of_step": prof[kname]["ms"] / sum(v["ms"] for v in prof.values()),
                    "all_kernels_ms": {k: v["ms"] / v["launches"] for k, v in prof.items()}}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": workload_config(args, B), "clocks": sampler.result(),
                "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s",
                        "h2d_bytes_per_step": B * SURF_ROWS * PITCH,
                        "d2h_bytes_per_step": B * (192 + DH * DW * 3),
                        "note": "pinned host NV12 surfaces -> H2D -> fused kernels -> D2H hist+resized"},
                "gpu_launches": int(launches), "roofline": roof,
                "whole_step_roofline_frac": value / world * B_ALG_FUSED / 1e9 / peak}
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def cpu_baseline(args):
    """Oracle port on a bounded sample, all host threads (one surface per thread)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    oracle.lib()
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(2 * cores, 16)
    surf = make_surfaces_np(min(sample, 16), 4321)

    def work(i):
        s = surf[i % len(surf)]
        oracle.nv12_hist_resize(s[:H], s[H:], DW, DH, W)

    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(work, range(cores)))
        t0 = time.perf_counter()
        list(ex.map(work, range(sample)))
        dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{sample} 1080p surfaces, oracle C port of NV12->RGB + Histogram + Resize(224)"}


if __name__ == "__main__":
    sys.exit(main())
