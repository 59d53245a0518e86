#!/usr/bin/env python
"""bench.py -- BASELINE.json configs[1]: 1080p decode + Resize(224) + Histogram, frames/s.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm (rank 0)

A "step" is one pass of the hot path over one batch of `--batch` decoded 1080p surfaces (NV12,
NVDEC layout, pitch 2048): surfaces -> RGB -> {Histogram 3x16 int32, Resize 224x224 RGB24}.
  value   frames/s over all ranks with the surfaces already resident in HBM (CUDA events, max
          over ranks); inputs rotate between two batches, each larger than the 126 MB L2.
  e2e     same metric end to end through the public pipeline API (scn_engine_run): HOST H.264
          byte streams in -> NVDEC -> Histogram + Resize(224) GPU ops -> result rows back on the
          host, every step.  h2d/d2h bytes are the encoded bytes fed and the rows returned.
  roofline  dominant kernel's algorithmic bytes / its CUDA-event duration (scn_prof_*), against
          MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the reference's CPU path restated with the libraries it calls (FFmpeg H.264
          decode through cv2.VideoCapture, cv2.calcHist x3, cv2.resize; one clip per process, as
          the reference runs one pipeline instance per core) on a bounded sample of the same
          clips on this box's host cores (rank 0, N=1 only).
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

    def setup(self):
        import pynvml as nv
        nv.nvmlInit()
        self.nv = nv
        self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        self.names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                      nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                      nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                      nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                      nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}

    def sample_now(self):
        """One sample from the calling thread (used while the timed steps are in flight)."""
        try:
            if not hasattr(self, "h"):
                self.setup()
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception as e:  # clocks are evidence, not a dependency
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        while not self.stop_flag:
            self.sample_now()
            time.sleep(0.005)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# DRAM bytes per launch (read + write) from the committed ncu --set full captures of the bench's own
# launch shape (64 surfaces 1920x1080, pitch 2048); see profiles/
NCU_DRAM_BYTES = {"nv12_hist_csa_kernel": 201362944 + 3201792}


def usable_cores():
    """Host cores this process may actually use: scheduler affinity capped by the cgroup CPU quota
    (the GPU boxes report 128 logical CPUs but run the container under cpu.max = 16 cores)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, -(-int(quota) // int(period))))
    except Exception:
        pass
    return max(1, n)


def make_surfaces_np(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 256, (n, SURF_ROWS, PITCH), dtype=np.uint8)
    return s


# ------------------------------------------------------------------------------------------
def make_clip_bytes(seed, frames, gop=30):
    """Synthetic 1080p H.264 (no encoder offline): I_PCM IDR every `gop` frames with
    uniform-random planes, the pictures in between P_Skip -> ~105 KB/frame, a realistic bitrate."""
    import numpy as np
    from scanner_b200 import engine as E
    rng = np.random.default_rng(seed)
    k = (frames + gop - 1) // gop
    yuv = rng.integers(0, 256, (k, H * W * 3 // 2), dtype=np.uint8)
    return E.h264_synth(yuv, W, H, gop=gop, non_key="skip", frames=frames)


def _ref_worker(path):
    """One reference pipeline instance: decode a clip and run Histogram + Resize on every frame
    (reference software_video_decoder.cpp:98-219 + tests/test_ops.cpp:13-59,114-170)."""
    import cv2
    cv2.setNumThreads(1)
    cap = cv2.VideoCapture(path)
    n = 0
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        for c in range(3):
            cv2.calcHist([frame], [c], None, [16], [0, 256])
        cv2.resize(frame, (DW, DH))
        n += 1
    return n


def cpu_reference_fps(clip_bytes, n_clips, steps, warmup):
    """frames/s of the CPU path on all host cores; clips are written to /dev/shm once."""
    import multiprocessing as mp
    import tempfile
    cores = usable_cores()
    tmpdir = tempfile.mkdtemp(prefix="scn_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    paths = []
    for i in range(n_clips):
        p = os.path.join(tmpdir, f"clip{i}.h264")
        with open(p, "wb") as f:
            f.write(clip_bytes[i % len(clip_bytes)])
        paths.append(p)
    ctx = mp.get_context("fork")
    try:
        with ctx.Pool(min(cores, n_clips)) as pool:
            for _ in range(warmup):
                pool.map(_ref_worker, paths[:min(cores, n_clips)])
            t0 = time.perf_counter()
            frames = 0
            for _ in range(steps):
                frames += sum(pool.map(_ref_worker, paths))
            dt = time.perf_counter() - t0
    finally:
        for p in paths:
            os.unlink(p)
        os.rmdir(tmpdir)
    return frames / dt, dt / steps, cores, frames // steps


def run_reference(args, rank, world):
    """CPU arm (rank 0 only): FFmpeg decode + OpenCV Histogram/Resize, one clip per core."""
    if rank != 0:
        return 0
    cores = usable_cores()
    frames_per_clip = 60
    n_clips = max(2 * cores, 8)
    clips = [make_clip_bytes(500 + i, frames_per_clip) for i in range(min(n_clips, 4))]
    fps, s_per_step, cores, sample = cpu_reference_fps(clips, n_clips, args.steps, max(1, min(args.warmup, 1)))
    desc = (f"{n_clips} clips x {frames_per_clip} frames per step ({sample} frames), cv2.VideoCapture (FFmpeg) decode + "
            "cv2.calcHist x3 + cv2.resize(224), one process per core")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, sample),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def workload_config(args, batch):
    return {"workload": "configs[1]: 1080p H.264 decode + Resize(224x224) + Histogram",
            "frame": [H, W], "pitch": PITCH, "resize": [DH, DW], "frames_per_step": batch,
            "value_leg": "decoded NV12 surfaces resident in HBM -> fused Histogram+Resize kernels "
                         "(two rotating input batches, each > the 126 MB L2)",
            "e2e_leg": "host H.264 (I_PCM IDR / 30 + P_Skip, ~105 KB/frame) -> NVDEC -> GPU ops -> host rows, "
                       "through scn_engine_run"}


# ------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  NCCL (and anything else native) prints banners such
    as "NCCL version ..." straight to fd 1, so fd 1 is pointed at stderr for the whole run and the
    result line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256,
                    help="surfaces per step (kernels take 64 per launch: a step of 256 is 4 launches per kernel)")
    ap.add_argument("--e2e-clips", type=int, default=56)
    ap.add_argument("--e2e-frames", type=int, default=120)
    ap.add_argument("--instances", type=int, default=0,
                    help="pipeline instances per GPU for the e2e leg (0 = 14 capped by 2 x usable host cores / ranks)")
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
    sampler.sample_now()
    sampler.samples.clear()
    sampler.start()
    l0 = L.scn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sampler.sample_now()  # the steps are still executing: a sample under load even for short runs
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.scn_launch_count() - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # --- roofline leg: the same kernels on the same batches with per-kernel events (scn_prof_*), not
    # used for `value`.  In the value leg the Resize kernel overlaps the Histogram kernel on a forked
    # stream; here each kernel is launched on its own (Histogram-only call, then Resize-only call) so
    # that a launch's duration is that kernel's and nothing else's.
    L.scn_prof_enable(1)
    for i in range(args.steps):
        kernels.nv12_hist_resize(batches[i & 1], W, H, DW, DH, plan, want_resize=False)
        kernels.nv12_hist_resize(batches[i & 1], W, H, DW, DH, plan, want_hist=False)
    torch.cuda.synchronize()
    prof = cabi.prof_report()
    L.scn_prof_enable(0)

    # --- e2e leg: the public pipeline, host H.264 in, host rows out (NVDEC + GPU ops + D2H)
    from scanner_b200 import engine as E
    from scanner_b200 import protolite
    E.load_stdlib()
    std = protolite.parse_proto(open(os.path.join(ROOT, "scanner_b200", "csrc", "ops", "stdlib_args.proto")).read())
    e2e_clips, e2e_frames = args.e2e_clips, args.e2e_frames
    if args.instances <= 0:
        # one decode session per NVDEC engine (7 on a B200) is the measured optimum -- 7: 8.3-8.8 K
        # frames/s, 8: 6.3 K, 14: 7.5-8.0 K, 21-42: 6.1-7.0 K (profiles/r01_e2e_engine_scaling.md): an
        # extra session shares an engine and the slowest pair sets the wall time -- but never more
        # threads than the rank's share of the host cores can keep fed
        engines = E.nvdec_caps(local_rank).get("engines", 7) or 7
        args.instances = max(2, min(engines, 2 * usable_cores() // world))
    eng = E.Engine(gpus=[local_rank], instances_per_gpu=args.instances)
    uniq = [make_clip_bytes(2000 + 16 * rank + i, e2e_frames) for i in range(min(e2e_clips, 4))]
    sids = [eng.add_h264(uniq[i % len(uniq)]) for i in range(e2e_clips)]
    graph = E.Graph()
    src = graph.add_source(True)
    op_h = graph.add_op("Histogram", [(src, "frame")], device=1)
    op_r = graph.add_op("Resize", [(src, "frame")], device=1)
    sink_h = graph.add_sink((op_h, "histogram"))
    sink_r = graph.add_sink((op_r, "frame"))
    jobs = []
    for sid in sids:
        j = E.Job()
        j.bind_source(src, sid)
        j.set_stream_args(op_r, protolite.encode(std["ResizeArgs"], {"width": DW, "height": DH}))
        jobs.append(j)
    for _ in range(2):
        eng.run(graph, jobs, 30, 60)  # warm-up: decoder creation, memory pools
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.run(graph, jobs, 30, 60)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_frames_total = e2e_clips * e2e_frames * args.steps
    e2e_stats = eng.stats()["counters"]
    assert jobs[0].output_rows(sink_h) == e2e_frames and jobs[0].output_rows(sink_r) == e2e_frames
    e2e_h2d = sum(len(uniq[i % len(uniq)]) for i in range(e2e_clips))
    e2e_d2h = e2e_clips * e2e_frames * (192 + DH * DW * 3)
    eng.close()

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
            per_launch = min(B, 64)  # SCN_MAX_PTRS surfaces per launch
            alg = (B_ALG_HIST_NV12 if kname.startswith("nv12_hist") else B_ALG_FUSED) * per_launch
            ach = alg / per_launch_s / 1e9
            roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "peak_kind": peak_kind + " (burst copy, MEASURED_PEAKS.json)",
                    "alg_bytes_per_launch": alg, "ms_per_launch": per_launch_s * 1e3,
                    "traffic": NCU_DRAM_BYTES.get(kname) if B % 64 == 0 else None,
                    "traffic_source": "dram__bytes_read+write of one ncu --set full capture of this launch "
                                      "shape (profiles/r01_nv12_hist_csa_v3.md)",
                    "launch_mode": "kernels timed one at a time; in the value leg Resize overlaps Histogram "
                                   "on a forked stream",
                    "kernel_share_of_step": prof[kname]["ms"] / sum(v["ms"] for v in prof.values()),
                    "all_kernels_ms": {k: v["ms"] / v["launches"] for k, v in prof.items()}}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": workload_config(args, B), "clocks": sampler.result(),
                "e2e": {"value": e2e_frames_total * world / (e2e_ms * 1e-3), "unit": "frames/s",
                        "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                        "frames_per_step": e2e_clips * e2e_frames, "pipeline_instances": args.instances,
                        "frames_decoded_last_step": e2e_stats.get("frames_decoded"),
                        "timing": "host wall clock around scn_engine_run (the call returns after the last "
                                  "row is on the host), max over ranks",
                        "bound": "NVDEC (7 engines/GPU): see profiles/r01_e2e_engine_scaling.md"},
                "gpu_launches": int(launches), "roofline": roof,
                "whole_step_roofline_frac": value / world * B_ALG_FUSED / 1e9 / peak}
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def cpu_baseline(args):
    """The reference's CPU path (FFmpeg decode + OpenCV ops) on a bounded sample, all host cores."""
    cores = usable_cores()
    n_clips = max(2 * cores, 8)
    clips = [make_clip_bytes(700 + i, 60) for i in range(min(n_clips, 4))]
    fps, s_per_step, cores, sample = cpu_reference_fps(clips, n_clips, 1, 1)
    return {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{sample} frames ({n_clips} clips x 60), cv2.VideoCapture (FFmpeg) decode + cv2.calcHist x3 + "
                      "cv2.resize(224), one process per core"}


if __name__ == "__main__":
    sys.exit(main())
