#!/usr/bin/env python3
"""Per-stage device times of encode_resident on the PHOTO workload (distinct 768x512 crops of the reference's
test photographs) beside the synthetic one, with the token rate."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import image_webp_b200 as Z
from image_webp_b200 import synth
import photo_inputs as PI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = Z.Context(0)
KEYS = ("yuv_ms", "analysis_ms", "pass1_ms", "chroma1_ms", "stats_ms", "chroma2_ms", "pass2_ms", "token_ms", "boolcode_ms", "assemble_ms")
for name, imgs in (("synthetic", list(synth.batch_photo_like(n, 768, 512, 0))), ("photo", list(PI.batch(n)))):
    ctx.stage(imgs)
    for q, m in ((75, 4), (75, 6), (50, 0)):
        p = Z.EncoderParams.lossy(q); p.method = m
        for _ in range(3):
            t = ctx.encode_resident(p)
        px = n * 768 * 512
        print("%s q%d m%d: total %.1f ms -> %.0f MPix/s; %.2f symbols/px; " % (name, q, m, t["device_total_ms"], px / t["device_total_ms"] / 1e3,
              t["symbols"] / px) + " ".join("%s %.1f" % (k[:-3], t[k]) for k in KEYS), flush=True)
    outs, _ = ctx.download()
    print("%s: %.3f B/px" % (name, sum(len(o) for o in outs) / (n * 768 * 512)), flush=True)
