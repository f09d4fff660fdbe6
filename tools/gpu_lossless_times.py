"""Lossless (VP8L) stage times on the GPU box: 1024 photo crops and 1024 synthetic images, 768x512 RGB."""
import json
import sys
import time

import numpy as np
import torch

import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import image_webp_b200 as Z
import photo_inputs as PI
from image_webp_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = Z.Context(0)
res = {}
for name in ("photo", "synthetic"):
    host = torch.empty((n, 512, 768, 3), dtype=torch.uint8).pin_memory().numpy()
    if name == "photo":
        PI.batch(n, out=host)
    else:
        for i in range(n):
            host[i] = synth.photo_like(768, 512, i)
    imgs = list(host)
    best = None
    for it in range(4):
        t0 = time.perf_counter()
        outs, t = ctx.encode_batch(imgs, Z.EncoderParams(), Z.ColorType.Rgb8)
        t["call_ms"] = (time.perf_counter() - t0) * 1e3
        if best is None or t["device_total_ms"] < best["device_total_ms"]:
            best = t
    best["bytes_per_px"] = sum(len(o) for o in outs) / (n * 768 * 512)
    best["kernel_mpix_s"] = n * 768 * 512 / best["device_total_ms"] / 1e3
    best["e2e_mpix_s"] = n * 768 * 512 / best["call_ms"] / 1e3
    res[name] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items() if v}
print(json.dumps(res, indent=1))
