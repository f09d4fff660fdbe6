#!/usr/bin/env python3
"""Device times of the on-device decoder / verifier (zw_verify) on the bench workloads: 1024 synthetic and 1024 photo
768x512 images, q75 m4.  usage: gpu_verify_times.py [n=1024]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import image_webp_b200 as Z
from image_webp_b200 import synth
import photo_inputs as PI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = Z.Context(0)
p = Z.EncoderParams.lossy(75); p.method = 4
for name, imgs in (("synthetic", list(synth.batch_photo_like(n, 768, 512, 0))), ("photo", list(PI.batch(n)))):
    for rep in range(2):
        pend = ctx.submit(imgs, p)
        info, ms = Z.verify_pending(pend)
        outs, t = pend.result()
    ps = np.array([i["psnr_rgb"] for i in info])
    bad = sum(1 for i in info if i["status"] != 0)
    px = n * 768 * 512
    print("%s: encode %.1f ms | verify: parse %.2f ms, reconstruct %.2f ms, filter %.2f ms, colour %.2f ms -> %.0f MPix/s decoded; psnr min %.2f mean %.2f; %d bad; %.2f symbols/px" % (
        name, t["device_total_ms"], ms[0], ms[1], ms[2], ms[3], px / sum(ms) / 1e3, ps.min(), ps.mean(), bad, t["symbols"] / px), flush=True)
