#!/usr/bin/env python3
"""Per-stage device times of encode_resident for a few (quality, method) settings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import image_webp_b200 as Z
from image_webp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
imgs = list(synth.batch_photo_like(n, 768, 512, 0))
ctx = Z.Context(0)
ctx.stage(imgs)
for q, m in ((75, 0), (75, 2), (75, 3), (75, 4), (75, 6)):
    p = Z.EncoderParams.lossy(q); p.method = m
    for _ in range(2):
        t = ctx.encode_resident(p)
    print("q%d m%d: total %.1f ms  p1 %.1f  st %.1f  p2 %.1f  tok %.1f  bc %.1f  -> %.0f MPix/s" % (
        q, m, t["device_total_ms"], t["pass1_ms"], t["stats_ms"], t["pass2_ms"], t["token_ms"], t["boolcode_ms"],
        n * 768 * 512 / t["device_total_ms"] / 1e3))
