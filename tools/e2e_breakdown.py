#!/usr/bin/env python
"""Where does the host time of the end-to-end path go?  56 clips x 120 frames 1080p of the realistic stream
(synth_h264) through the engine with 7 instances: profiler interval totals (summed over instances) per run, in
memory and with the save stage writing tables.  usage: python tools/e2e_breakdown.py [cavlc|pcm]"""
import json
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from scanner_b200 import engine as E  # noqa: E402
from scanner_b200 import protolite, synth_h264  # noqa: E402

STD = protolite.parse_proto(open("scanner_b200/csrc/ops/stdlib_args.proto").read())
kind = sys.argv[1] if len(sys.argv) > 1 else "cavlc"


def clip(seed, n):
    if kind == "cavlc":
        return synth_h264.write(1920, 1080, n, gop=30, seed=seed)[0]
    rng = np.random.default_rng(seed)
    yuv = rng.integers(0, 256, ((n + 29) // 30, 1080 * 1920 * 3 // 2), dtype=np.uint8)
    return E.h264_synth(yuv, 1920, 1080, gop=30, non_key="skip", frames=n)


E.load_stdlib()
uniq = [clip(2000 + i, 120) for i in range(4)]
root = tempfile.mkdtemp(dir="/dev/shm")
db = E.Database(root)
for i in range(56):
    db.ingest_h264(f"clip_{i}", uniq[i % 4])
eng = E.Engine(gpus=[0], instances_per_gpu=7)
sids = [db.add_video_stream(eng, f"clip_{i}") for i in range(56)]
g = E.Graph()
src = g.add_source(True)
hs = g.add_op("Histogram", [(src, "frame")], device=1)
rz = g.add_op("Resize", [(src, "frame")], device=1)
sh, sr = g.add_sink((hs, "histogram")), g.add_sink((rz, "frame"))
for mode in ("memory", "tables", "hist_only_memory"):
    for run in range(3):
        jobs, tabs = [], []
        for i, s in enumerate(sids):
            j = E.Job()
            j.bind_source(src, s)
            j.set_stream_args(rz, protolite.encode(STD["ResizeArgs"], {"width": 224, "height": 224}))
            if mode == "tables":
                th = db.new_table(f"h_{run}_{i}", "histogram", False, "Histogram", i)
                tr = db.new_table(f"r_{run}_{i}", "frame", True, "", i)
                j.set_sink_table(sh, th, keep_rows=False)
                j.set_sink_table(sr, tr, keep_rows=False)
                tabs.append((th, tr))
            jobs.append(j)
        t0 = time.time()
        if mode == "hist_only_memory":
            g2 = E.Graph()
            s2 = g2.add_source(True)
            h2 = g2.add_op("Histogram", [(s2, "frame")], device=1)
            g2.add_sink((h2, "histogram"))
            jobs = []
            for s in sids:
                j = E.Job()
                j.bind_source(s2, s)
                jobs.append(j)
            eng.run(g2, jobs, 30, 60)
        else:
            eng.run(g, jobs, 30, 60, out_dir=root if mode == "tables" else None)
            for j, (th, tr) in zip(jobs, tabs):
                db.commit_job_table(th, j)
                db.commit_job_table(tr, j)
        dt = time.time() - t0
        st = eng.stats()
        if run == 2:
            iv = {k: round(v) for k, v in sorted(st["intervals_ms"].items(), key=lambda kv: -kv[1])[:12]}
            c = st["counters"]
            print(json.dumps({"stream": kind, "mode": mode, "fps": round(56 * 120 / dt), "wall_ms": round(dt * 1e3),
                              "intervals_ms_summed_over_7_instances": iv,
                              "nvdec_us": {k: v for k, v in c.items() if k.startswith("nvdec_")},
                              "session_rate": round(c["inst0_frames_decoded"] * 1e6 / max(1, c["inst0_decode_busy_us"]))}), flush=True)
        if mode == "tables":
            for i in range(56):
                db.delete_table(f"h_{run}_{i}")
                db.delete_table(f"r_{run}_{i}")
eng.close()
shutil.rmtree(root, ignore_errors=True)
