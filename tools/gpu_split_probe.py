#!/usr/bin/env python3
"""Blocking batch call (zw_encode_webp_batch) under different chunk splits: config 2 and config 4."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import image_webp_b200 as Z
from image_webp_b200 import synth
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True); t.numpy()[...] = a; return t.numpy()
ctx = Z.Context(0)
c2 = pinned(synth.batch_photo_like(1024, 768, 512, 0))
base = pinned(np.stack([synth.photo_like(1920, 1080, 100 + i) for i in range(8)]))
for name, imgs, q, m in (("config2", [c2[i] for i in range(1024)], 75, 4), ("config4", [base[i % 8] for i in range(256)], 75, 6)):
    p = Z.EncoderParams.lossy(q); p.method = m
    prep = ctx.prepare(imgs)
    for split in ("", "1", "2"):
        if split: os.environ["ZW_SPLIT"] = split
        else: os.environ.pop("ZW_SPLIT", None)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); zo, rc = ctx.encode_batch_raw(prep, p); dt = time.perf_counter() - t0
            ctx._collect(zo, True); best = min(best, dt)
        print("%s split=%s: %.1f ms" % (name, split or "auto", 1e3 * best), flush=True)
