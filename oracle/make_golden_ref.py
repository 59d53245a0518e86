#!/usr/bin/env python3
"""Golden vectors from the REFERENCE'S OWN kernel (oracle/_ref/libref_image.so = scanner/util/image.cu
compiled unmodified).  Runs on a GPU box:

    gpurun -- 'python oracle/make_golden_ref.py'      -> gpurun_out/nv12_ref.npz
    cp gpurun_out/nv12_ref.npz tests/golden/

Contents: the RGB24 output of scanner::convertNV12toRGBA for the seeded surfaces of nv12_np.npz
(same seeds/shapes, oracle/synth.py), and the SHA-256 of its output over the exhaustive input set
synth.nv12_exhaustive (every (Y,Cb,Cr) triple on an even row + averaged-chroma odd rows).  The CPU
suite (tests/test_oracle_golden.py) holds the oracle to both; the GPU suite re-runs the kernel itself.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(300, 16, 32, 32), (301, 270, 480, 512), (302, 1080, 1920, 2048), (303, 34, 70, 128), (304, 2, 2, 64),
         (305, 18, 30, 32), (306, 64, 62, 64)]


def main():
    import torch

    from oracle import ref, synth
    out = {}
    for i, (seed, h, w, pitch) in enumerate(CASES):
        luma, chroma = synth.nv12_surface(seed, h, w, pitch)
        surf = torch.from_numpy(np.concatenate([luma, chroma], 0)).cuda()
        out[f"c{i}_meta"] = np.array([seed, h, w, pitch], np.int64)
        out[f"c{i}_out"] = ref.nv12_to_rgb(surf, w, h).cpu().numpy()
    sha = hashlib.sha256()
    for f in range(synth.EXH_FRAMES):
        luma, chroma = synth.nv12_exhaustive(f)
        surf = torch.from_numpy(np.concatenate([luma, chroma], 0)).cuda()
        sha.update(ref.nv12_to_rgb(surf, synth.EXH_W, synth.EXH_H).cpu().numpy().tobytes())
    out["exhaustive_sha256"] = np.frombuffer(sha.digest(), np.uint8)
    out["provenance"] = np.array(
        f"scanner::convertNV12toRGBA from /root/reference/scanner/util/image.cu, nvcc default flags, sm_100a, "
        f"run on {torch.cuda.get_device_name(0)}")
    dst = os.path.join(ROOT, "gpurun_out", "nv12_ref.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, sha.hexdigest())


if __name__ == "__main__":
    main()
