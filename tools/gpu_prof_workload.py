#!/usr/bin/env python3
"""The bench workload once, for ncu: stage n synthetic (or photo) 768x512 images, `reps` x encode_resident at q75 m4.
usage: gpu_prof_workload.py [n=1024] [synthetic|photo] [reps=1]   (ZW_QUAD=0/1/2 selects the luma kernels)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import image_webp_b200 as Z
from image_webp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kind = sys.argv[2] if len(sys.argv) > 2 else "synthetic"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
if kind == "photo":
    import photo_inputs as PI
    imgs = list(PI.batch(n))
else:
    imgs = list(synth.batch_photo_like(n, 768, 512, 0))
ctx = Z.Context(0)
ctx.stage(imgs)
p = Z.EncoderParams.lossy(75); p.method = 4
for _ in range(reps):
    t = ctx.encode_resident(p)
if os.environ.get("ZW_HASH"):
    import hashlib
    outs, _ = ctx.download()
    print("hash", hashlib.sha256(b"".join(hashlib.sha256(o).digest() for o in outs)).hexdigest()[:16])
print("%s n=%d: total %.1f ms pass1 %.1f pass2 %.1f launches %d" % (kind, n, t["device_total_ms"], t["pass1_ms"], t["pass2_ms"], t["kernel_launches"]))
