"""The reference's OWN op library inside this engine.

oracle/_ref/libref_test_ops.so is /root/reference/tests/test_ops.cpp compiled UNMODIFIED against this repo's plugin
headers (oracle/Makefile; the cv:: calls of its Histogram / Resize / OpticalFlow kernels are stubs that abort -- there
is no OpenCV C++ here).  Two things follow from loading it with scn_load_op_library:
  * the drop-in boundary is real: the reference's REGISTER_OP / REGISTER_KERNEL source registers and runs here;
  * the kernels whose arithmetic is written out in that file are the reference itself: its Blur pins oracle.blur (and
    with it every GPU blur kernel, which the -m gpu tests hold to the oracle), its TestIncrement* ops pin the engine's
    bounded-state warm-up and unbounded-state semantics with the values the reference's tests/py_test.py expects.
The library registers the names Histogram / Resize / Blur / ... too, so every check runs in its own process.
Skipped where the library was not built (no /root/reference)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_test_ops.so")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_test_ops.so not built")

PRELUDE = r'''
import os, struct, sys
import numpy as np
sys.path.insert(0, os.environ["SCN_ROOT"])
import oracle
from oracle import synth
from scanner_b200 import engine as E, protolite
E.load_op_library(os.path.join(os.environ["SCN_ROOT"], "oracle", "_ref", "libref_test_ops.so"))
REF_ARGS = protolite.parse_proto(open("/root/reference/tests/test_ops.proto").read()) if os.path.exists(
    "/root/reference/tests/test_ops.proto") else protolite.parse_proto(
    "message BlurArgs { int32 kernel_size = 1; float sigma = 2; }")
'''


def run(body):
    out = subprocess.run([sys.executable, "-c", PRELUDE + body], env=dict(os.environ, SCN_ROOT=ROOT), cwd=ROOT,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "REF_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
    return out.stdout


def test_the_reference_op_source_registers_through_the_plugin_headers():
    out = run(r'''
ops = E.list_ops()
for name in ("Histogram", "OpticalFlow", "Resize", "Blur", "TestIncrementUnbounded", "TestIncrementUnboundedFrame",
             "TestIncrementBounded", "TestIncrementBoundedFrame", "Sleep", "SleepFrame"):
    assert name in ops, name
assert ops["Blur"]["protobuf_name"] == "BlurArgs" and ops["Resize"]["stream_protobuf_name"] == "ResizeArgs"
assert ops["OpticalFlow"]["can_stencil"] and ops["TestIncrementBounded"]["bounded"] and ops["TestIncrementUnbounded"]["unbounded"]
print("REF_OK")
''')
    assert "REF_OK" in out


def test_reference_blur_kernel_pins_the_oracle():
    """reference tests/test_ops.cpp:239-310 (BlurKernel, CPU) run by this engine == oracle.blur on the region the
    reference writes (it leaves the border of its new frame uninitialised), odd and even kernel sizes."""
    run(r'''
for (k, h, w, n) in [(3, 37, 53, 3), (4, 40, 31, 2), (5, 64, 48, 2), (8, 33, 47, 1), (15, 40, 56, 1)]:
    frames = np.stack([synth.rand_frame(900 + 7 * k + i, h, w) for i in range(n)])
    eng = E.Engine(gpus=[], cpu_instances=2)
    g = E.Graph(); src = g.add_source(True)
    bl = g.add_op("Blur", [(src, "frame")], device=0,
                  args=protolite.encode(REF_ARGS["BlurArgs"], {"kernel_size": k, "sigma": 0.5}))
    sink = g.add_sink((bl, "frame"))
    j = E.Job(); j.bind_source(src, eng.add_raw_frames(frames))
    eng.run(g, [j], 2, 4)
    ys, xs = oracle.blur_interior(h, w, k)
    for i in range(n):
        got = j.output_row(sink, i)
        assert got.shape == (h, w, 3)
        assert (got[ys, xs] == oracle.blur(frames[i], k)[ys, xs]).all(), (k, i)
    eng.close()
print("REF_OK")
''')


def test_reference_increment_ops_give_the_values_the_reference_tests_expect():
    """py_test.py:407-423 (bounded state, warm-up 3, Gather [0,10,25,26,27] -> [0,3,3,4,5]) and :426-435 (unbounded
    state under Slice(all(50)) / Unslice: as many rows as the input; here also the values: a count that restarts
    with every slice group) with the reference's own TestIncrementKernel (:173-236)."""
    run(r'''
n = 120
rows = [struct.pack("<q", i) for i in range(n)]
eng = E.Engine(gpus=[], cpu_instances=2)
# bounded
g = E.Graph(); src = g.add_source(False)
inc = g.add_op("TestIncrementBounded", [(src, "column")], warmup=3)
s = g.add_sample((inc, "integer")); sink = g.add_sink((s, "integer"))
j = E.Job(); j.bind_source(src, eng.add_bytes(rows))
j.set_sampler(s, "Gather", protolite.encode(protolite.SAMPLER_ARGS["GatherSamplerArgs"], {"rows": [0, 10, 25, 26, 27]}))
eng.run(g, [j], 10, 100)
assert [struct.unpack("<q", j.output_row(sink, i))[0] for i in range(5)] == [0, 3, 3, 4, 5]
# unbounded under Slice / Unslice
g = E.Graph(); src = g.add_source(False)
sl = g.add_slice((src, "column"))
inc = g.add_op("TestIncrementUnbounded", [(sl, "column")])
un = g.add_unslice((inc, "integer")); sink = g.add_sink((un, "integer"))
j = E.Job(); j.bind_source(src, eng.add_bytes(rows))
j.set_partitioner(sl, "Strided", protolite.encode(protolite.SAMPLER_ARGS["StridedPartitionerArgs"], {"stride": 1, "group_size": 50}))
eng.run(g, [j], 5, 25)
got = [struct.unpack("<q", j.output_row(sink, i))[0] for i in range(j.output_rows(sink))]
assert len(got) == n and got == [i % 50 for i in range(n)], got[:60]
eng.close()
print("REF_OK")
''')
