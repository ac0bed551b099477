"""Seeded synthetic image + depth frames shared by tests, bench.py and
the golden-fixture generator.  numpy PCG64 streams only, so every machine sees the same bytes.

Depth classes exercise different failure modes of the warp:
  'scene'   smooth ramp + 3 discs + 1% noise (the benchmark input)
  'noise'   white noise (worst-case occlusion)
  'quant'   the scene quantised to a few 8-bit levels (exact closeness ties, Q7)
  'flat'    constant depth (max == min -> zeros)
  'steps'   vertical bars at a handful of levels (hard edges for the blur)
  'card'    the reference's own create_test_images.py layout (3 discs on a vertical ramp)
"""
import numpy as np


def make_image(n, h, w, seed=0, black_box=False):
    rng = np.random.default_rng(1000 + seed)
    img = rng.random((n, h, w, 3), dtype=np.float32)
    if black_box:  # genuinely black source pixels (mask quirk Q6)
        img[:, h // 4:h // 2, w // 4:w // 2, :] = 0.0
    return img


def _discs(h, w, shift=0.0):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    d = 0.25 + 0.2 * xx / w + 0.1 * yy / h
    for cx, cy, r, v in ((0.30, 0.35, 0.15, 0.6), (0.65, 0.55, 0.20, 0.8), (0.50, 0.80, 0.08, 1.0)):
        m = (xx - (cx + shift) * w) ** 2 + (yy - cy * h) ** 2 <= (r * h) ** 2
        d = np.where(m, np.float32(v), d)
    return d.astype(np.float32)


def make_depth(n, h, w, kind="scene", seed=0, channels=3, scale255=False):
    rng = np.random.default_rng(2000 + seed)
    out = np.empty((n, h, w), np.float32)
    for i in range(n):
        shift = 0.002 * i  # discs drift frame to frame (video)
        if kind == "scene":
            d = _discs(h, w, shift) + 0.01 * rng.random((h, w), dtype=np.float32)
        elif kind == "noise":
            d = rng.random((h, w), dtype=np.float32)
        elif kind == "quant":
            d = np.round(_discs(h, w, shift) * 12.0) / 12.0
            d = np.round(d * 255.0) / 255.0
        elif kind == "flat":
            d = np.full((h, w), 0.5, np.float32)
        elif kind == "steps":
            levels = np.array([0.1, 0.9, 0.3, 0.7, 0.2, 1.0, 0.0, 0.5], np.float32)
            d = np.broadcast_to(levels[(np.arange(w) * 8 // max(w, 1)) % 8], (h, w)).copy()
            d[h // 3: 2 * h // 3, w // 5: 3 * w // 5] = 0.95
        elif kind == "card":
            yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
            d = (80.0 + 50.0 * yy / h) / 255.0
            for cx, cy, r, v in ((0.25, 0.5, 0.166, 100), (0.5, 0.5, 0.2, 170), (0.75, 0.5, 0.133, 240)):
                d = np.where((xx - cx * w) ** 2 + (yy - cy * h) ** 2 <= (r * h) ** 2, np.float32(v / 255.0), d)
            d = np.floor(d * 255.0) / 255.0
        else:
            raise ValueError(kind)
        out[i] = np.clip(d, 0.0, 1.0)
    if scale255:
        out = out * np.float32(255.0)
    if channels == 1:
        return out[..., None].copy()
    return np.repeat(out[..., None], 3, axis=-1)


def index_probe_image(h, w):
    """uint8-exact image encoding (column + 1) in R,G and a constant in B:
    after any uint8 CPU fill, R + 256*G - 1 is the source column, 0 means 'unfilled'."""
    col = np.arange(w, dtype=np.int64) + 1
    img = np.zeros((h, w, 3), np.uint8)
    img[..., 0] = (col & 255)[None, :]
    img[..., 1] = (col >> 8)[None, :]
    img[..., 2] = 77
    return img


def probe_to_float(img_u8):
    """float image whose truncating quantisation (Q2) gives back img_u8 exactly."""
    return ((img_u8.astype(np.float32) + np.float32(0.5)) / np.float32(255.0)).astype(np.float32)


def dark_case(seed, h=24, w=333):
    """Dark image + noisy depth (0..255) that exercise the interpolating fill's rare paths: ramps between black or
    near-black borders, black ramp values that start new gaps, black-but-filled pixels (SIG:1871-1892)."""
    rng = np.random.default_rng(50 + seed)
    img = rng.integers(0, 3, (h, w, 3), dtype=np.uint8) * rng.integers(0, 2, (h, w, 1), dtype=np.uint8)
    img[:, ::7] = rng.integers(0, 256, (h, (w + 6) // 7, 3), dtype=np.uint8)
    if seed % 3 == 2:
        img[:] = 0
        img[:, 100:110] = 9
    d = (rng.random((h, w), dtype=np.float32) * np.float32(255)).astype(np.float32)
    if seed % 3:
        d[:, 150:200] = np.float32(200)
    return img, d


def fuzz_case(rng):
    """One random small stage case (image uint8 [h,w,3], depth 0..255 [h,w], divergence, separation, exponent,
    convergence) of the kind oracle/fuzz_vs_reference.py and tests/test_gpu_fuzz.py throw at every fill."""
    h = int(rng.integers(1, 20))
    w = int(rng.integers(2, 200))
    style = int(rng.integers(0, 5))
    if style == 0:      # uniform noise
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif style == 1:    # dark: ramps between near-black borders, black pixels that are really there
        img = rng.integers(0, 3, (h, w, 3), dtype=np.uint8) * rng.integers(0, 2, (h, w, 1), dtype=np.uint8)
        img[:, ::7] = rng.integers(0, 256, (h, (w + 6) // 7, 3), dtype=np.uint8)
    elif style == 2:    # mostly black
        img = np.zeros((h, w, 3), np.uint8)
        img[:, w // 3: w // 3 + 4] = rng.integers(0, 256, 3, dtype=np.uint8)
    elif style == 3:    # bright (uint8 sums wrap)
        img = rng.integers(250, 256, (h, w, 3), dtype=np.uint8)
    else:               # smooth gradient
        img = np.broadcast_to((np.arange(w) * 255 // max(w - 1, 1)).astype(np.uint8)[None, :, None], (h, w, 3)).copy()
    dstyle = int(rng.integers(0, 5))
    if dstyle == 0:
        d = rng.random((h, w), dtype=np.float32) * np.float32(255)
    elif dstyle == 1:
        d = np.broadcast_to(np.linspace(0, 255, w, dtype=np.float32)[None], (h, w)).copy()
    elif dstyle == 2:
        d = np.full((h, w), 128, np.float32)
        d[:, w // 4: w // 2] = 250
    elif dstyle == 3:
        d = np.round(rng.random((h, w), dtype=np.float32) * 4) * np.float32(60)
    else:
        d = np.full((h, w), 77, np.float32)   # flat
    div = float(rng.choice([0.5, 2.0, 3.5, 6.0, 10.0, 15.0])) * float(rng.choice([-1, 1]))
    sep = float(rng.choice([0.0, 0.0, 1.0, -2.5]))
    expo = float(rng.choice([1.0, 2.0, 0.7]))
    conv = float(rng.choice([0.0, 0.5, 1.0, 0.3]))
    return img, d.astype(np.float32), div, sep, expo, conv
