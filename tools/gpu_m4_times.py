#!/usr/bin/env python3
"""Per-stage device times of the bench workload (q75 m4) for the library in ZW_LIB_PATH (or the default)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_webp_b200 as Z
from image_webp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
q = int(sys.argv[2]) if len(sys.argv) > 2 else 75
m = int(sys.argv[3]) if len(sys.argv) > 3 else 4
imgs = list(synth.batch_photo_like(n, 768, 512, 0))
ctx = Z.Context(0)
ctx.stage(imgs)
p = Z.EncoderParams.lossy(q); p.method = m
for _ in range(3):
    t = ctx.encode_resident(p)
print("%s q%d m%d n%d: total %.1f ms  p1 %.1f c1 %.1f st %.1f c2 %.1f p2 %.1f tok %.1f bc %.1f -> %.0f MPix/s" % (
    os.environ.get("ZW_LIB_PATH", "default"), q, m, n, t["device_total_ms"], t["pass1_ms"], t["chroma1_ms"], t["stats_ms"], t["chroma2_ms"],
    t["pass2_ms"], t["token_ms"], t["boolcode_ms"], n * 768 * 512 / t["device_total_ms"] / 1e3))
