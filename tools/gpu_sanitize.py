#!/usr/bin/env python3
"""Small mixed workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import image_webp_b200 as Z
from image_webp_b200 import synth
import oracle_lib as O
ctx = Z.Context(0)
imgs = [synth.photo_like(320, 272, 1), synth.photo_like(99, 87, 2), synth.photo_like(64, 48, 3), synth.noise(48, 80, 4), synth.photo_like(16, 16, 5)]
for q, m in ((75, 4), (50, 0), (75, 6)):
    p = Z.EncoderParams.lossy(q); p.method = m
    outs, _ = ctx.encode_batch(imgs, p)
    for im, o in zip(imgs, outs):
        rc, ref, _ = O.encode(im, q, m)
        assert o == ref
g = imgs[0][:, :, 1:2].copy()
outs, _ = ctx.encode_batch([g], Z.EncoderParams.lossy(75), color=Z.ColorType.L8)
print("sanitize workload ok")
