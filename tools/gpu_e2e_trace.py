#!/usr/bin/env python3
"""Timeline of bench.py's end-to-end leg (BatchPipeline, depth 3, views): when every submit / result returns."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import image_webp_b200 as Z
from image_webp_b200 import synth
n, steps = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 10
host = torch.empty((n, 512, 768, 3), dtype=torch.uint8, pin_memory=True)
host.numpy()[...] = synth.batch_photo_like(n, 768, 512, 0)
imgs = [host.numpy()[i] for i in range(n)]
p = Z.EncoderParams.lossy(75); p.method = 4
if len(sys.argv) > 2:  # like bench.py: a second context that ran the kernel-only leg first
    ctx = Z.Context(0)
    ctx.stage(imgs)
    for _ in range(5):
        ctx.encode_resident(p)
    if sys.argv[2] == "download":
        ctx.download()
pipe = Z.BatchPipeline(0, depth=3, views=True)
prep = pipe.ctx.prepare(imgs)
for f in [pipe.submit(prep, p) for _ in range(3)]:
    f.result()
torch.cuda.synchronize()
t0 = time.perf_counter()
ev = []
futs = []
for s in range(steps):
    futs.append(pipe.submit(prep, p)); ev.append(("submit %d" % s, time.perf_counter() - t0))
for k, f in enumerate(futs):
    outs, t = f.result(); ev.append(("result %d dev %.1f h2d %.1f d2h %.1f" % (k, t["device_total_ms"], t["h2d_ms"], t["d2h_ms"]), time.perf_counter() - t0))
torch.cuda.synchronize()
wall = time.perf_counter() - t0
for name, t in ev:
    print("%8.1f ms  %s" % (1e3 * t, name))
print("wall %.1f ms -> %.2f ms/step" % (1e3 * wall, 1e3 * wall / steps))
