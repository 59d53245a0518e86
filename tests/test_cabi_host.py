"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/scn_kernels.h
declares, and its host-only entry points (resize target rule, resize plan) agree with the
oracle.  No compute launches here -- there is no GPU in the build container."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from oracle import synth
from scanner_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"SCN_API\s+[\w\s\*]+?\b(scn_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols("scn_kernels.h")
    assert len(names) >= 13
    l = ctypes.CDLL(cabi.LIB_PATH)
    for n in names:
        assert hasattr(l, n), f"libscn_kernels.so does not export {n}"
    # and the Python binding table covers the header exactly
    assert sorted(cabi.SIGNATURES) == names


def test_abi_version_and_launch_counter():
    l = cabi.lib()
    assert l.scn_abi_version() == 1
    assert l.scn_launch_count() == 0  # nothing launched on a CPU-only run


def test_resize_target_matches_oracle():
    l = cabi.lib()
    for (sw, sh, w, h, mn, pa) in [(1920, 1080, 224, 224, 0, 0), (1920, 1080, 0, 540, 0, 1),
                                   (1920, 1080, 640, 0, 0, 1), (640, 480, 1000, 1000, 1, 0),
                                   (640, 480, 320, 1000, 1, 0), (640, 480, 0, 0, 0, 0),
                                   (33, 77, 100, 0, 1, 1)]:
        ow, oh = ctypes.c_int(), ctypes.c_int()
        l.scn_resize_target(sw, sh, w, h, mn, pa, ctypes.byref(ow), ctypes.byref(oh))
        assert (ow.value, oh.value) == oracle.resize_target(sw, sh, w, h, bool(mn), bool(pa))


def _apply_plan_numpy(img, plan_bytes, dw, dh):
    """Apply the PRODUCT's plan with numpy -- checks the host table logic against the oracle."""
    p = plan_bytes.view(np.int32)
    hdr, taps = p[:8], p[8:].reshape(-1, 4)
    assert hdr[3] == dw and hdr[4] == dh
    xt, yt = taps[:dw], taps[dw:dw + dh]
    s = img.astype(np.int32)
    hb = s[:, xt[:, 0], :] * xt[None, :, 2, None] + s[:, xt[:, 1], :] * xt[None, :, 3, None]
    h0, h1 = hb[yt[:, 0]], hb[yt[:, 1]]
    b0, b1 = yt[:, 2][:, None, None], yt[:, 3][:, None, None]
    return ((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


@pytest.mark.parametrize("h,w,dh,dw", [(48, 64, 20, 30), (100, 130, 333, 257), (7, 9, 224, 224),
                                       (1080, 1920, 224, 224), (37, 53, 37, 53), (64, 64, 1, 1),
                                       (5, 1, 3, 4)])
def test_resize_plan_matches_oracle(h, w, dh, dw):
    l = cabi.lib()
    n = l.scn_resize_plan_bytes(dw, dh)
    assert n == 32 + 16 * (dw + dh)
    buf = np.zeros(n, np.uint8)
    assert l.scn_resize_plan_fill(w, h, dw, dh, buf.ctypes.data) == 0
    img = synth.rand_frame(7, h, w)
    assert (_apply_plan_numpy(img, buf, dw, dh) == oracle.resize(img, dw, dh)).all()


def test_bad_arguments_are_rejected_without_a_gpu():
    l = cabi.lib()
    assert l.scn_resize_plan_fill(0, 10, 4, 4, None) == -1
    assert l.scn_resize_plan_bytes(0, 5) == 0
    assert l.scn_hist16_u8c3(None, -1, 4, 4, None, None) == -1
    assert l.scn_hist16_u8c3(None, 0, 4, 4, None, None) == 0  # n == 0 is a no-op
    assert l.scn_box_blur_u8c3(None, 1, 4, 4, 0, None, None) == -1  # kernel_size < 1
    assert l.scn_nv12_to_rgb24(None, None, 8, 1, 7, 8, None, 24, None) == -1  # odd width


def test_no_silent_cpu_fallback():
    """On a box without CUDA the torch wrappers must raise, not compute on the CPU."""
    import torch
    from scanner_b200 import kernels
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cabi.ScnError):
        kernels.histogram(torch.zeros((1, 4, 4, 3), dtype=torch.uint8))


def test_engine_library_exports_every_declared_symbol():
    import re as _re
    from scanner_b200 import engine as E
    src = open(os.path.join(ROOT, "include", "scn_engine.h")).read()
    names = sorted(set(_re.findall(r"SCN_ENGINE_API\s+[\w\s\*]+?\b(scn_\w+)\s*\(", src)))
    assert len(names) >= 30
    l = ctypes.CDLL(E.ENGINE_PATH)
    for n in names:
        assert hasattr(l, n), f"libscn_engine.so does not export {n}"
    assert sorted(E.SIGNATURES) == names
    # the stdlib plugin loads and registers the three GPU ops without touching a GPU
    E.load_stdlib()
    ops = E.list_ops()
    for op in ("Histogram", "Resize", "Blur"):
        assert op in ops and E.lib().scn_kernel_registered(op.encode(), 1) == 1
