#!/usr/bin/env python
"""Exhaustive check: can the NV12 -> RGB -> 16-bin histogram path use cheaper arithmetic than the
reference's float colour matrix and still be bit-exact?

Only floor(clamp(v, 0, 1023) / 64) of each channel reaches the histogram, so ANY arithmetic that
yields the same bin for every input is exact by exhaustion: R and B depend on (Y, Cr) / (Y, Cb)
-- 2^16 inputs each --, G on all 2^24.  Candidate (one FMA per pixel and channel, both clamps from
`.sat`, no separate upper clamp):

    x = fma.sat(Y, a, c)         a = 4*1.1644/992,  c = per-chroma-sample constant (1-2 FMAs, shared by 4 pixels)
    m = floor(15.5 * x)          fma.rm(x, 15.5, 2^23); 15.5/992 == 1/64, and x == 1.0 (v >= 992) gives 15

Result on this machine (numpy, FMA emulated exactly in long double):
    R: 0 mismatches of 65,536      B: 0 of 65,536      (reference values come no closer than 1.6e-3 to a bin edge)
    G: 46-53 mismatches of 16,777,216 for every constant nudge tried (reference values DO land on bin edges)
=> R and B can take the cheap form, G needs the reference's exact chain.  See DESIGN.md section 10.

    python tools/search_bin_arithmetic.py [--g]      (--g adds the 2^24 sweep, ~25 min)
"""
import sys

import numpy as np

f32, LD = np.float32, np.longdouble


def fma(a, b, c):
    """correctly rounded float32 fma: a*b+c is exact in long double for these magnitudes"""
    return (a.astype(LD) * b.astype(LD) + c.astype(LD)).astype(np.float32)


def nudge(v, ulps):
    for _ in range(abs(ulps)):
        v = np.nextafter(v, f32(np.inf) if ulps > 0 else f32(-np.inf))
    return v


def full(v):
    return np.full((256, 256), v, dtype=np.float32)


def sweep_rb():
    y4 = np.broadcast_to(np.arange(256, dtype=np.float32)[:, None] * f32(4), (256, 256)).copy()
    c4 = np.broadcast_to(np.arange(256, dtype=np.float32)[None, :] * f32(4) - f32(512), (256, 256)).copy()
    yb = np.broadcast_to(np.arange(256, dtype=np.float32)[:, None], (256, 256)).copy()
    cb = np.broadcast_to(np.arange(256, dtype=np.float32)[None, :] - f32(128), (256, 256)).copy()
    for name, k in (("R", 1.596), ("B", 2.0172)):
        ref = fma(c4, full(f32(k)), (y4 * f32(1.1644)).astype(np.float32))          # image.cu, FMA-contracted
        want = (np.clip(ref, 0, 1023).astype(np.uint32) >> 6).astype(np.int32)
        inside = ref[(ref > 0) & (ref < 1023)] % 64
        print(f"{name}: closest reference value to a bin edge: {np.minimum(inside, 64 - inside).min():.2e}")
        for da, dc in ((0, 0), (1, 1), (-1, -1), (3, -3)):
            a, kc = nudge(f32(4 * 1.1644 / 992), da), nudge(f32(4 * k / 992), dc)
            x = np.clip(fma(yb, full(a), (cb * kc).astype(np.float32)), 0, 1)
            m = np.floor(x.astype(LD) * LD(15.5)).astype(np.int32)
            print(f"   constants nudged by ({da:+d}, {dc:+d}) ulp: {int((m != want).sum())} mismatches of 65536")


def sweep_g():
    cb4 = np.broadcast_to(np.arange(256, dtype=np.float32)[:, None] * f32(4) - f32(512), (256, 256)).copy()
    cr4 = np.broadcast_to(np.arange(256, dtype=np.float32)[None, :] * f32(4) - f32(512), (256, 256)).copy()
    cbb = np.broadcast_to(np.arange(256, dtype=np.float32)[:, None] - f32(128), (256, 256)).copy()
    crb = np.broadcast_to(np.arange(256, dtype=np.float32)[None, :] - f32(128), (256, 256)).copy()
    k0, k1, k2 = f32(1.1644), f32(-0.3918), f32(-0.813)
    variants = [(da, d1, d2) for da in (-1, 0, 1) for d1 in (-1, 0, 1) for d2 in (-1, 0, 1)]
    consts = {}
    for _, d1, d2 in variants:
        if (d1, d2) not in consts:
            t = (cbb * nudge(f32(4 * -0.3918 / 992), d1)).astype(np.float32)
            consts[(d1, d2)] = fma(crb, full(nudge(f32(4 * -0.813 / 992), d2)), t)
    bad = dict.fromkeys(variants, 0)
    for y in range(256):
        ly = full(f32(y * 4) * k0)
        g = fma(cr4, full(k2), fma(cb4, full(k1), ly))
        want = (np.clip(g, 0, 1023).astype(np.uint32) >> 6).astype(np.int32)
        for v in variants:
            x = np.clip(fma(full(f32(y)), full(nudge(f32(4 * 1.1644 / 992), v[0])), consts[(v[1], v[2])]), 0, 1)
            bad[v] += int((np.floor(x.astype(LD) * LD(15.5)).astype(np.int32) != want).sum())
    best = sorted(bad.items(), key=lambda kv: kv[1])
    print("G: mismatches of 16777216 per (a, k1, k2) nudge, best three:", best[:3], "worst:", best[-1])


if __name__ == "__main__":
    sweep_rb()
    if "--g" in sys.argv:
        sweep_g()
