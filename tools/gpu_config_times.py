#!/usr/bin/env python3
"""Kernel-only device times of the other BASELINE.json configurations (3, 4, 5) on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import image_webp_b200 as Z
from image_webp_b200 import synth
ctx = Z.Context(0)
def run(name, imgs, q, m):
    ctx.stage(imgs)
    p = Z.EncoderParams.lossy(q); p.method = m
    for _ in range(2):
        t = ctx.encode_resident(p)
    px = sum(i.shape[0] * i.shape[1] for i in imgs)
    print("%s: total %.1f ms  p1 %.1f c1 %.1f st %.1f c2 %.1f p2 %.1f tok %.1f bc %.1f -> %.0f MPix/s" % (
        name, t["device_total_ms"], t["pass1_ms"], t["chroma1_ms"], t["stats_ms"], t["chroma2_ms"], t["pass2_ms"], t["token_ms"],
        t["boolcode_ms"], px / t["device_total_ms"] / 1e3), flush=True)
run("config1 1 x 768x512 q75 m4", [synth.photo_like(768, 512, 0)], 75, 4)
run("config3 1 x 4096x4096 q90 m4", [synth.photo_like(4096, 4096, 3, freq_scale=4.0)], 90, 4)
base = [synth.photo_like(1920, 1080, 100 + i) for i in range(8)]
run("config4 256 x 1920x1080 q75 m6", [base[i % 8] for i in range(256)], 75, 6)
tb = [synth.photo_like(256, 256, 200 + i) for i in range(64)]
if len(sys.argv) < 2:
    run("config5 (1/8 shard) 8192 x 256x256 q50 m0", [tb[i % 64] for i in range(8192)], 50, 0)
