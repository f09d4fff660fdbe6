#!/usr/bin/env python3
"""e2e timing breakdown of encode_batch (pinned host buffers) for a few lane counts / batch sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import image_webp_b200 as Z
from image_webp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
host = torch.empty((n, 512, 768, 3), dtype=torch.uint8, pin_memory=True)
host.numpy()[...] = synth.batch_photo_like(n, 768, 512, 0)
imgs = [host.numpy()[i] for i in range(n)]
ctx = Z.Context(0, lanes=lanes)
p = Z.EncoderParams.lossy(75); p.method = 4
for _ in range(2):
    ctx.encode_batch(imgs, p)
t0 = time.perf_counter()
outs, t = ctx.encode_batch(imgs, p)
wall = (time.perf_counter() - t0) * 1e3
print("n%d lanes%d: wall %.1f ms (lib wall %.1f) device %.1f h2d %.1f d2h %.1f  -> e2e %.0f MPix/s, kernel %.0f MPix/s" % (
    n, lanes, wall, t["wall_ms"], t["device_total_ms"], t["h2d_ms"], t["d2h_ms"], n * 768 * 512 / wall / 1e3, n * 768 * 512 / t["device_total_ms"] / 1e3))
