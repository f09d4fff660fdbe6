"""Deterministic synthetic RGB inputs: a "photo-like" generator G(seed) that
yields a mix of I16 / I4 / skipped macroblocks and several populated segments, plus the
edge-case inputs the parity tests use.  numpy only; no file or network access."""
import numpy as np


def _value_noise(h, w, cell, rng):
    gh, gw = h // cell + 2, w // cell + 2
    g = rng.random((gh, gw), dtype=np.float32)
    ys = np.arange(h, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0 = ys.astype(np.int32)
    x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy)
    fx = fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def photo_like(width, height, seed=0, freq_scale=1.0):
    """G(seed): smooth gradients + two octaves of value noise + a few hard edges + flat patches."""
    rng = np.random.default_rng(0x5EED0000 + int(seed))
    y = np.arange(height, dtype=np.float32)[:, None]
    x = np.arange(width, dtype=np.float32)[None, :]
    out = np.empty((height, width, 3), dtype=np.float32)
    n8 = _value_noise(height, width, max(2, int(8 * freq_scale)), rng) - 0.5
    n2 = _value_noise(height, width, max(1, int(2 * freq_scale)), rng) - 0.5
    n32 = _value_noise(height, width, max(4, int(32 * freq_scale)), rng) - 0.5
    tex_mask = (_value_noise(height, width, max(8, int(48 * freq_scale)), rng) > 0.55).astype(np.float32)
    # hard-edged rectangles / flat areas
    edge = np.zeros((height, width), dtype=np.float32)
    for _ in range(6):
        x0 = int(rng.integers(0, max(1, width - 8)))
        y0 = int(rng.integers(0, max(1, height - 8)))
        ww = int(rng.integers(8, max(9, width // 3)))
        hh = int(rng.integers(8, max(9, height // 3)))
        edge[y0:y0 + hh, x0:x0 + ww] += float(rng.integers(-60, 60))
    flat = _value_noise(height, width, max(8, int(64 * freq_scale)), rng) > 0.72
    for c in range(3):
        phi = 2.1 * c + 0.37 * (seed % 7)
        v = (128 + 64 * np.sin(2 * np.pi * x / (197 * freq_scale) + phi)
             + 48 * np.sin(2 * np.pi * y / (131 * freq_scale) + 0.5 * phi)
             + 90 * n32 + (32 * n8 + 24 * n2 * tex_mask) + edge * (0.6 + 0.2 * c))
        v = np.where(flat, 96.0 + 20 * c, v)
        out[:, :, c] = v
    return np.clip(out + 0.5, 0, 255).astype(np.uint8)


def gradient(width, height):
    """64x64-gradient-style image used by the reference's quality tests
    (tests/lossy_encoder_quality.rs:163-174 shape)."""
    y = np.arange(height, dtype=np.uint32)[:, None]
    x = np.arange(width, dtype=np.uint32)[None, :]
    r = (x * 255 // max(1, width - 1)).astype(np.uint8) + np.zeros((height, 1), np.uint8)
    g = (y * 255 // max(1, height - 1)).astype(np.uint8) + np.zeros((1, width), np.uint8)
    b = np.full((height, width), 128, np.uint8)
    return np.ascontiguousarray(np.stack([r, g, b], axis=2))


def checker_gradient(width, height, cell=16):
    y = np.arange(height)[:, None]
    x = np.arange(width)[None, :]
    chk = (((x // cell) + (y // cell)) & 1).astype(np.uint8)
    r = (chk * 200 + 20).astype(np.uint8)
    g = ((x * 255 // max(1, width - 1)) + 0 * y).astype(np.uint8)
    b = ((y * 255 // max(1, height - 1)) + 0 * x).astype(np.uint8)
    return np.ascontiguousarray(np.stack([r, g, b], axis=2))


def solid(width, height, rgb=(120, 130, 140)):
    return np.ascontiguousarray(np.broadcast_to(np.array(rgb, np.uint8), (height, width, 3)))


def noise(width, height, seed=1):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (height, width, 3), dtype=np.uint8)


def batch_photo_like(n, width, height, seed0=0):
    """n images G(seed0+i) stacked [n,h,w,3]; distinct seeds cycle every 64 images with a
    per-image brightness/phase tweak so large batches stay cheap to generate."""
    base = [photo_like(width, height, seed0 + i) for i in range(min(n, 64))]
    out = np.empty((n, height, width, 3), np.uint8)
    for i in range(n):
        b = base[i % len(base)]
        k = i // len(base)
        if k == 0:
            out[i] = b
        else:
            out[i] = np.roll(b, (7 * k) % height, axis=0) ^ np.uint8(k & 3)
    return out
