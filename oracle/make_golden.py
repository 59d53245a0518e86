#!/usr/bin/env python
"""Generate tests/golden/*.npz -- run in the BUILD container, outputs are committed.

Pins for the oracle (oracle/scn_oracle.c):
  * Histogram / Resize: produced by the library that holds the reference's arithmetic,
    OpenCV (`cv2.calcHist`, `cv2.resize`; this container has opencv-python-headless 4.13.0,
    the reference pins 4.2.0 in deps.sh:643) called exactly as tests/test_ops.cpp:38-43,156
    call it.  The reference itself ships no golden vectors for these ops.
  * Blur / NV12->RGB: the reference's arithmetic is fully in-tree (tests/test_ops.cpp:265-294,
    scanner/util/image.cu:67-200) and cannot be compiled here, so the pin is a second,
    independent numpy restatement written in this file (vectorised, float32 with explicit
    fused multiply-adds emulated in float64 where exact).
Inputs are stored by seed (numpy PCG64 `default_rng(seed).integers`) when large, verbatim when
small.  Nothing here is imported by the product.
"""
import math
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.synth import rand_frame, smooth_frame, nv12_surface, flow_pair  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def cv_hist(img):
    return np.stack([cv2.calcHist([img], [j], None, [16], [0, 256]).astype(np.int32).ravel()
                     for j in range(3)])


def np_blur(img, k):
    fl = int(math.ceil(k / 2.0)) - 1
    fr = k // 2
    h, w, _ = img.shape
    out = np.zeros_like(img)
    if h - fr <= fl or w - fr <= fl:
        return out
    acc = np.zeros((h - fl - fr, w - fl - fr, 3), np.uint32)
    for ry in range(-fl, fr + 1):
        for rx in range(-fl, fr + 1):
            acc += img[fl + ry:h - fr + ry, fl + rx:w - fr + rx].astype(np.uint32)
    out[fl:h - fr, fl:w - fr] = (acc // ((fl + fr + 1) ** 2)).astype(np.uint8)
    return out


def _fma32(a, b, c):
    """float32 fma: exact product+sum in float64 then one rounding (|values| < 2^24 so the
    float64 intermediate is exact: 24-bit x 24-bit products fit 53 bits, sums stay exact)."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def np_nv12_to_rgb(luma, chroma, width):
    h, pitch = luma.shape
    ys = np.arange(h)
    yc = ys >> 1
    cbcr = chroma.astype(np.uint32)
    nxt = np.minimum(yc + 1, h // 2 - 1)
    avg = (cbcr[yc] + cbcr[nxt] + 1) >> 1
    use_avg = ((ys & 1) == 1) & (yc < (h >> 1) - 1)
    c = np.where(use_avg[:, None], avg, cbcr[yc])
    xs = np.arange(width) & ~1
    cb = c[:, xs].astype(np.int32)
    cr = c[:, xs + 1].astype(np.int32)
    l = (luma[:, :width].astype(np.uint32) << 2).astype(np.float32)
    fcb = ((cb << 2) - 512).astype(np.float32)
    fcr = ((cr << 2) - 512).astype(np.float32)
    # nvcc's contraction of image.cu:81-90 (SASS of oracle/_ref): fma(cr,k2, fma(y,k0, fl(cb*k1)))
    k0 = np.float32(1.1644)
    r = _fma32(fcr, np.float32(1.596), _fma32(l, k0, (fcb * np.float32(0.0)).astype(np.float32)))
    g = _fma32(fcr, np.float32(-0.813), _fma32(l, k0, (fcb * np.float32(-0.3918)).astype(np.float32)))
    b = _fma32(fcr, np.float32(0.0), _fma32(l, k0, (fcb * np.float32(2.0172)).astype(np.float32)))
    out = np.stack([r, g, b], -1)
    out = np.minimum(np.maximum(out, np.float32(0)), np.float32(1023))
    return (out.astype(np.uint32) >> 2).astype(np.uint8)


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- histogram
    cases = {}
    for i, (h, w, kind) in enumerate([(48, 64, "rand"), (480, 640, "smooth"), (1, 1, "rand"),
                                      (17, 31, "rand"), (1080, 1920, "rand"), (33, 5, "smooth")]):
        seed = 100 + i
        img = rand_frame(seed, h, w) if kind == "rand" else smooth_frame(seed, h, w)
        cases[f"c{i}_meta"] = np.array([seed, h, w, 0 if kind == "rand" else 1], np.int64)
        cases[f"c{i}_hist"] = cv_hist(img)
    # degenerate: all one value
    cases["const_meta"] = np.array([255, 64, 64, 2], np.int64)
    cases["const_hist"] = cv_hist(np.full((64, 64, 3), 255, np.uint8))
    np.savez_compressed(os.path.join(OUT, "hist_cv2.npz"), **cases)

    # ---- resize
    cases = {}
    for i, (h, w, dh, dw, kind) in enumerate([
            (48, 64, 24, 32, "rand"),      # exact 2x -> area fast path
            (48, 64, 20, 30, "rand"),
            (100, 130, 333, 257, "rand"),  # upscale (top/bottom rows clamp both taps)
            (7, 9, 224, 224, "rand"),
            (480, 640, 224, 224, "smooth"),
            (1080, 1920, 224, 224, "rand"),
            (37, 53, 37, 53, "rand"),      # identity
            (64, 64, 1, 1, "rand"),
            (480, 640, 100, 640, "rand")]):
        seed = 200 + i
        img = rand_frame(seed, h, w) if kind == "rand" else smooth_frame(seed, h, w)
        cases[f"c{i}_meta"] = np.array([seed, h, w, dh, dw, 0 if kind == "rand" else 1], np.int64)
        cases[f"c{i}_out"] = cv2.resize(img, (dw, dh))
    np.savez_compressed(os.path.join(OUT, "resize_cv2.npz"), **cases)

    # ---- blur (independent numpy restatement)
    cases = {}
    for i, (h, w, k) in enumerate([(48, 64, 3), (48, 64, 5), (31, 17, 4), (20, 20, 1), (9, 9, 9),
                                   (5, 40, 7), (240, 320, 3)]):
        seed = 300 + i
        img = rand_frame(seed, h, w)
        cases[f"c{i}_meta"] = np.array([seed, h, w, k], np.int64)
        cases[f"c{i}_out"] = np_blur(img, k)
    np.savez_compressed(os.path.join(OUT, "blur_np.npz"), **cases)

    # ---- nv12 -> rgb (independent numpy restatement)
    cases = {}
    for i, (h, w, pitch) in enumerate([(16, 32, 32), (48, 64, 128), (2, 2, 16), (270, 480, 512),
                                       (30, 50, 64)]):
        seed = 400 + i
        luma, chroma = nv12_surface(seed, h, w, pitch)
        cases[f"c{i}_meta"] = np.array([seed, h, w, pitch], np.int64)
        cases[f"c{i}_out"] = np_nv12_to_rgb(luma, chroma, w)
    np.savez_compressed(os.path.join(OUT, "nv12_np.npz"), **cases)
    # ---- optical flow: cv2 gray conversion + cv2 Farneback with the reference's parameters
    cases = {}
    for i, (h, w, kind) in enumerate([(96, 128, "shift"), (64, 80, "shift"), (120, 160, "noise"), (37, 53, "shift")]):
        seed = 500 + i
        a, b = flow_pair(seed, h, w, kind)
        cases[f"c{i}_meta"] = np.array([seed, h, w, 0 if kind == "shift" else 1], np.int64)
        g0, g1 = cv2.cvtColor(a, cv2.COLOR_BGR2GRAY), cv2.cvtColor(b, cv2.COLOR_BGR2GRAY)
        cases[f"c{i}_gray0"] = g0
        cases[f"c{i}_flow"] = cv2.FarnebackOpticalFlow_create(3, 0.5, False, 15, 3, 5, 1.2, 0).calc(g0, g1, None)
    np.savez_compressed(os.path.join(OUT, "flow_cv2.npz"), **cases)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())
