"""N>1 path on the CPU: two gloo ranks shard clips between them, each runs its own engine, rank 0
gathers and checks every row against the oracle.  (The GPU run uses the same code with nccl.)"""
import os
import subprocess
import sys

import pytest

from scanner_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, struct
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["SCN_ROOT"])
import oracle
from oracle import synth
from scanner_b200 import engine as E, shard

dist.init_process_group("gloo")
rank, world = shard.rank_world()
assert (rank, world) == (dist.get_rank(), dist.get_world_size())
E.load_op_library(os.path.join(os.environ["SCN_ROOT"], "build", "tests", "libtest_plugin_ops.so"))
n_clips = 7
lengths = [5 + 3 * i for i in range(n_clips)]
mine = shard.shard_indices(n_clips, rank, world, weights=lengths)
eng = E.Engine(gpus=[], cpu_instances=2)
g = E.Graph(); src = g.add_source(True)
h = g.add_op("TestHistogramOracle", [(src, "frame")]); sink = g.add_sink((h, "histogram"))
jobs, frames = {}, {}
for i in mine:
    frames[i] = np.stack([synth.rand_frame(1000 * i + k, 24, 32) for k in range(lengths[i])])
    j = E.Job(); j.bind_source(src, eng.add_raw_frames(frames[i])); jobs[i] = j
eng.run(g, [jobs[i] for i in mine], 4, 8)
local = {i: jobs[i].output_array(sink, 192, np.int32).reshape(-1, 3, 16) for i in mine}
allrows = shard.gather_rows(local, n_clips)
dist.barrier()
if rank == 0:
    assert all(r is not None for r in allrows)
    for i in range(n_clips):
        fr = np.stack([synth.rand_frame(1000 * i + k, 24, 32) for k in range(lengths[i])])
        assert allrows[i].shape == (lengths[i], 3, 16)
        for k in range(lengths[i]):
            assert (allrows[i][k] == oracle.hist16(fr[k])).all()
    print("MULTIRANK_OK", sorted(mine), [len(a) for a in allrows])
dist.destroy_process_group()
'''


def test_shard_indices_partition_and_balance():
    for world in (1, 2, 3, 8):
        owned = [shard.shard_indices(10, r, world) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(10))
    w = [100, 1, 1, 1, 1, 50, 50]
    parts = [shard.shard_indices(len(w), r, 2, weights=w) for r in range(2)]
    assert sorted(sum(parts, [])) == list(range(len(w)))
    loads = [sum(w[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 2
    assert shard.gather_rows({0: "a", 1: "b"}, 2) == ["a", "b"]  # no process group: local passthrough


def test_two_gloo_ranks_shard_clips(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liborc.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SCN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", str(script)]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "MULTIRANK_OK" in out.stdout, out.stdout[-3000:]


HALO_WORKER = r'''
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.environ["SCN_ROOT"])
from scanner_b200 import halo
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 23
seq = torch.arange(n * 6, dtype=torch.float32).reshape(n, 2, 3)       # the whole "clip"
a, b = halo.interval_of(n, rank, world)
mine = seq[a:b].clone()
for stencil in ([0, 1], [-1, 0, 1], [-2, 0], [0]):
    wins = halo.stencil_windows(mine, stencil)
    for k, s in enumerate(sorted(stencil)):
        idx = torch.clamp(torch.arange(a, b) + s, 0, n - 1)
        assert torch.equal(wins[k], seq[idx]), (rank, stencil, s)
# a stencil op computed shard-wise equals the single-process result
cur, nxt = halo.stencil_windows(mine, [0, 1])
diff = (nxt - cur).abs().sum(dim=(1, 2))
full = (seq[torch.clamp(torch.arange(n) + 1, max=n - 1)] - seq).abs().sum(dim=(1, 2))
assert torch.equal(diff, full[a:b])
dist.barrier()
if rank == 0:
    print("HALO_OK", world)
dist.destroy_process_group()
'''


def test_interval_of_partitions():
    from scanner_b200 import halo
    for n in (1, 7, 23, 64):
        for world in (1, 2, 3, 4, 8):
            parts = [halo.interval_of(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(tmp_path, world):
    script = tmp_path / "halo_worker.py"
    script.write_text(HALO_WORKER)
    env = dict(os.environ, SCN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29620 + world), str(script)]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:]
    assert f"HALO_OK {world}" in out.stdout, out.stdout[-3000:]


ENGINE_HALO_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["SCN_ROOT"])
from oracle import synth
from scanner_b200 import engine as E, halo
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
E.load_op_library(os.path.join(os.environ["SCN_ROOT"], "build", "tests", "libtest_plugin_ops.so"))
lengths = [23, 9, 40]                                   # three clips, each split over all ranks
clips = [np.stack([synth.rand_frame(100 * c + k, 12, 16) for k in range(n)]) for c, n in enumerate(lengths)]
eng = E.Engine(gpus=[], cpu_instances=2)
eng.init_host_halo()
g = E.Graph(); src = g.add_source(True)
d = g.add_op("TestFrameDiff", [(src, "frame")]); sink = g.add_sink((d, "diff"))
jobs, windows = [], []
for c, n in enumerate(lengths):
    bounds = [halo.interval_of(n, r, world)[0] for r in range(world)] + [n]
    a, b = bounds[rank], bounds[rank + 1]
    mine = clips[c].copy()
    mine[:a] = 0; mine[b:] = 0                          # rows of other ranks are NOT available here:
    j = E.Job(); j.bind_source(src, eng.add_raw_frames(mine))   # a correct result proves they arrived over the wire
    j.set_shard(rank, bounds, list(range(world)))
    jobs.append(j); windows.append((a, b))
eng.run(g, jobs, 2, 4)
st = eng.stats()["counters"]
for c, n in enumerate(lengths):
    a, b = windows[c]
    assert jobs[c].output_rows(sink) == b - a
    got = jobs[c].output_array(sink, 8, np.float64, row0=a).ravel()
    nxt = clips[c][np.minimum(np.arange(a, b) + 1, n - 1)].astype(np.int64)
    want = np.abs(nxt - clips[c][a:b].astype(np.int64)).reshape(b - a, -1).sum(1).astype(np.float64)
    assert (got == want).all(), (rank, c, got, want)
frame = 12 * 16 * 3
expect_recv = sum(frame for c, n in enumerate(lengths) if windows[c][1] < n and windows[c][1] > windows[c][0])
assert st["halo_bytes_received"] == expect_recv, (rank, st["halo_bytes_received"], expect_recv)
dist.barrier()
if rank == 0:
    print("ENGINE_HALO_OK", world, st["halo_bytes_sent"], st["halo_bytes_received"])
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_engine_halo_exchange_for_sharded_clips_gloo(tmp_path, world):
    """configs[3] on the CPU: every rank computes one interval of every clip through scn_engine_run; the frame
    its stencil {0,1} needs from the next interval comes from the neighbouring rank (host transport over gloo),
    never from the rank's own (zeroed) copy."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    script = tmp_path / "engine_halo_worker.py"
    script.write_text(ENGINE_HALO_WORKER)
    env = dict(os.environ, SCN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29630 + world), str(script)]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:]
    assert f"ENGINE_HALO_OK {world}" in out.stdout, out.stdout[-3000:]


QUEUE_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["SCN_ROOT"])
import oracle
from oracle import synth
from scanner_b200 import engine as E

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
E.load_op_library(os.path.join(os.environ["SCN_ROOT"], "build", "tests", "libtest_plugin_ops.so"))
root = os.environ["SCN_TEST_DB"]
db = E.Database(root)
n_clips, lengths = 6, [9, 14, 5, 11, 8, 13]
frames = [np.stack([synth.rand_frame(3000 * i + k, 24, 32) for k in range(lengths[i])]) for i in range(n_clips)]
eng = E.Engine(gpus=[], cpu_instances=2)
g = E.Graph(); src = g.add_source(True)
h = g.add_op("TestHistogramOracle", [(src, "frame")]); sink = g.add_sink((h, "histogram"))
eng.share_task_queue(os.path.join(root, "task_queue"))
for run in range(2):                                   # the queue is reset between runs
    box = [db.new_tables([(f"h{run}_{i}", "histogram", False, "Histogram", i) for i in range(n_clips)]) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ids = box[0]
    jobs = []
    for i in range(n_clips):                           # EVERY rank lists every clip, in the same order
        j = E.Job(); j.bind_source(src, eng.add_raw_frames(frames[i]))
        j.set_sink_table(sink, ids[i], keep_rows=False)
        jobs.append(j)
    if rank == 0:
        eng.reset_task_queue()
    dist.barrier()
    eng.run(g, jobs, 2, 4, root)
    dist.barrier()
    done = [None] * world
    dist.all_gather_object(done, sum(v for k, v in eng.stats()["counters"].items() if k.startswith("inst") and k.endswith("_tasks")))
    if rank == 0:
        db.commit_job_tables(list(zip(ids, jobs)))
        total_tasks = sum((n + 3) // 4 for n in lengths)
        assert sum(done) == total_tasks, (done, total_tasks)          # every task ran exactly once, somewhere
        for i in range(n_clips):
            rows = db.read_rows(f"h{run}_{i}", "histogram")
            assert len(rows) == lengths[i]
            for k in range(lengths[i]):
                assert rows[k] == oracle.hist16(frames[i][k]).tobytes(), (run, i, k)
        print("QUEUE_OK", run, done)
    dist.barrier()
eng.close()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_pull_tasks_from_one_shared_queue_gloo(tmp_path, world):
    """scn_engine_share_task_queue: every rank lists the same jobs; each task of a run is executed by exactly one rank
    (the reference's workers pull tasks from the master), the items of all ranks land in the same tables."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liborc.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    script = tmp_path / "worker.py"
    script.write_text(QUEUE_WORKER)
    db = tmp_path / "db"
    env = dict(os.environ, SCN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT, SCN_TEST_DB=str(db))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29640 + world), str(script)]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:]
    assert out.stdout.count("QUEUE_OK") == 2, out.stdout[-3000:]
