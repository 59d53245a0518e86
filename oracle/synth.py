"""Seeded synthetic inputs shared by oracle/make_golden.py and tests/ (test infrastructure)."""
import numpy as np


def rand_frame(seed, h, w, c=3):
    return np.random.default_rng(seed).integers(0, 256, (h, w, c), dtype=np.uint8)


def smooth_frame(seed, h, w):
    """Config-1 style content: gradient + inverted 16x16 block + gaussian noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)),
                     ((xx + yy) * 255 // max(w + h - 2, 1))], -1).astype(np.float32)
    bx, by = (seed * 7) % max(w - 16, 1), (seed * 3) % max(h - 16, 1)
    base[by:by + 16, bx:bx + 16] = 255 - base[by:by + 16, bx:bx + 16]
    base += rng.normal(0, 8, base.shape)
    return np.clip(base, 0, 255).astype(np.uint8)


def frame(seed, h, w, kind):
    return rand_frame(seed, h, w) if kind == 0 else smooth_frame(seed, h, w)


def nv12_surface(seed, h, w, pitch=None):
    pitch = pitch or w
    rng = np.random.default_rng(seed)
    luma = rng.integers(0, 256, (h, pitch), dtype=np.uint8)
    chroma = rng.integers(0, 256, (h // 2, pitch), dtype=np.uint8)
    return luma, chroma
