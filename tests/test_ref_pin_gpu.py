"""Pin of the NV12 -> RGB arithmetic to the REFERENCE'S OWN KERNEL.

oracle/_ref/libref_image.so is /root/reference/scanner/util/image.cu compiled unmodified
(oracle/Makefile: `nvcc -DHAVE_CUDA -I/root/reference -gencode arch=compute_100a,code=sm_100a -shared`).
Here scanner::convertNV12toRGBA (image.cu:229-239 -> NV12_to_RGB :109-200) runs on the B200 over an
input set that presents every one of the 2^24 (Y,Cb,Cr) triples on an even row, the averaged-chroma
odd rows and every frame's last-chroma-row case (oracle/synth.py nv12_exhaustive), and its output must
equal, byte for byte,
  * the oracle restatement (oracle.nv12_to_rgb),
  * the product's colour conversion (scn_nv12_to_rgb24),
and the histogram / resize the product computes straight from the NV12 surface (scn_nv12_hist_resize:
nv12_hist kernels + nv12_resize_kernel) must equal the histogram / resize of the reference kernel's RGB.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from oracle import ref, synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reflib():
    if not ref.available():
        pytest.fail("oracle/_ref/libref_image.so is missing: run `make -C oracle ref` where /root/reference exists "
                    "(the built .so travels to the GPU box)")
    return ref


def _hist_of_rgb(rgb):  # (H,W,3) u8 cuda -> (3,16) int64, bin = v >> 4 (tests/test_ops.cpp:38-43)
    bins = (rgb >> 4).to(torch.int64).reshape(-1, 3)
    return torch.stack([torch.bincount(bins[:, c], minlength=16) for c in range(3)])


def test_reference_kernel_equals_oracle_and_product_on_all_yuv_triples(reflib):
    from scanner_b200 import kernels
    w, h = synth.EXH_W, synth.EXH_H
    seen = torch.zeros(1 << 24, dtype=torch.bool, device="cuda")
    for f in range(synth.EXH_FRAMES):
        luma, chroma = synth.nv12_exhaustive(f)
        surf_h = np.concatenate([luma, chroma], 0)
        surf = torch.from_numpy(surf_h).cuda()
        want = reflib.nv12_to_rgb(surf, w, h)
        # coverage bookkeeping: the triples presented on even rows of this frame
        ly = surf[0:h:2, :w].to(torch.int64)
        cb = surf[h:, 0:w:2].to(torch.int64).repeat_interleave(2, dim=1)
        cr = surf[h:, 1:w:2].to(torch.int64).repeat_interleave(2, dim=1)
        seen[(ly << 16 | cb << 8 | cr).reshape(-1)] = True
        # 1. the oracle restatement
        got_o = oracle.nv12_to_rgb(luma, chroma, w)
        bad = (want.cpu().numpy() != got_o)
        assert not bad.any(), f"frame {f}: oracle differs from the reference kernel at {bad.sum()} bytes, first {np.argwhere(bad)[0]}"
        # 2. the product's conversion kernel
        got_p = kernels.nv12_to_rgb(surf[None], w, h)[0]
        assert torch.equal(got_p, want), f"frame {f}: scn_nv12_to_rgb24 differs from the reference kernel"
        # 3. histogram + resize computed from the NV12 surface without materialising RGB
        hist, res = kernels.nv12_hist_resize(surf[None], w, h, 224, 224)
        assert torch.equal(hist[0].to(torch.int64), _hist_of_rgb(want)), f"frame {f}: fused histogram"
        assert torch.equal(res[0], kernels.resize(want[None], 224, 224)[0]), f"frame {f}: fused resize"
    assert bool(seen.all()), "the input set must present every (Y,Cb,Cr) triple on an even row"


@pytest.mark.parametrize("seed,h,w,pitch", [(1, 1080, 1920, 2048), (2, 270, 480, 512), (3, 34, 70, 128), (4, 2, 2, 64),
                                            (5, 18, 30, 32), (6, 1088, 1920, 1920)])
def test_reference_kernel_on_random_surfaces(reflib, seed, h, w, pitch):
    from scanner_b200 import kernels
    luma, chroma = synth.nv12_surface(seed, h, w, pitch)
    surf = torch.from_numpy(np.concatenate([luma, chroma], 0)).cuda()
    want = reflib.nv12_to_rgb(surf, w, h)
    assert (want.cpu().numpy() == oracle.nv12_to_rgb(luma, chroma, w)).all()
    assert torch.equal(kernels.nv12_to_rgb(surf[None], w, h)[0], want)
    hist, _ = kernels.nv12_hist_resize(surf[None], w, h, 0, 0, want_resize=False)
    assert torch.equal(hist[0].to(torch.int64), _hist_of_rgb(want))


def test_committed_reference_goldens_reproduce(reflib, golden_dir):
    """tests/golden/nv12_ref.npz was written by oracle/make_golden_ref.py from this same kernel."""
    path = os.path.join(golden_dir, "nv12_ref.npz")
    if not os.path.exists(path):
        pytest.skip("nv12_ref.npz not generated yet")
    g = np.load(path)
    for k in sorted(x[:-5] for x in g.files if x.endswith("_meta")):
        seed, h, w, pitch = [int(x) for x in g[k + "_meta"]]
        luma, chroma = synth.nv12_surface(seed, h, w, pitch)
        surf = torch.from_numpy(np.concatenate([luma, chroma], 0)).cuda()
        assert (reflib.nv12_to_rgb(surf, w, h).cpu().numpy() == g[k + "_out"]).all(), k
