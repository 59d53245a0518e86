"""The reference's own tutorial FILES, executed unmodified (where the reference tree is mounted).

tests/test_tutorials_gpu.py replays the tutorials' bodies on a GPU; this runs the actual scripts of
/root/reference/examples/tutorials that need no GPU op (01, 02, 04: Python kernels, batching,
stencils, bounded state, Slice / Unslice, save_mp4) through tests/helpers/run_reference_tutorial.py,
each in its own process (they register ops by global name).  Skipped where the tree is absent (the
GPU box)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TUTORIALS = "/root/reference/examples/tutorials"

pytestmark = pytest.mark.skipif(not os.path.isdir(TUTORIALS), reason="reference tree not mounted")


@pytest.mark.parametrize("script,frames,videos", [
    ("01_defining_python_ops.py", 60, ["01_resized_class.mp4", "01_resized_fn.mp4"]),
    ("02_op_attributes.py", 120, ["02_batch_resize.mp4", "02_device_resize.mp4", "02_flow.mp4", "02_masked.mp4"]),
    ("04_slicing.py", 1400, ["04_masked.mp4"]),  # its scenes are rows 1100-1400
])
def test_reference_tutorial_runs_unmodified(script, frames, videos):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "run_reference_tutorial.py"),
                          os.path.join(TUTORIALS, script), str(frames)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "Finished!" in out.stdout
    wrote = dict(re.findall(r"wrote (\S+) (\d+)", out.stdout))
    assert sorted(wrote) == videos and all(int(v) > 1000 for v in wrote.values())
