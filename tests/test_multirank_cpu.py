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
