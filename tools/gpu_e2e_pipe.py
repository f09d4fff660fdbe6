#!/usr/bin/env python3
"""e2e throughput with D contexts on one GPU driven by D host threads (double buffering across batches)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import image_webp_b200 as Z
from image_webp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
host = torch.empty((n, 512, 768, 3), dtype=torch.uint8, pin_memory=True)
host.numpy()[...] = synth.batch_photo_like(n, 768, 512, 0)
imgs = [host.numpy()[i] for i in range(n)]
p = Z.EncoderParams.lossy(75); p.method = 4
ctxs = [Z.Context(0) for _ in range(depth)]
for c in ctxs:
    c.encode_batch(imgs, p)
res = [None] * steps
def worker(k):
    for s in range(k, steps, depth):
        res[s] = ctxs[k].encode_batch(imgs, p)[0]
t0 = time.perf_counter()
th = [threading.Thread(target=worker, args=(k,)) for k in range(depth)]
for t in th: t.start()
for t in th: t.join()
wall = time.perf_counter() - t0
ok = all(r == res[0] for r in res)
print("n%d depth%d steps%d: %.1f ms/step -> e2e %.0f MPix/s, outputs identical across steps: %s" % (n, depth, steps, 1e3 * wall / steps, steps * n * 768 * 512 / wall / 1e6, ok))
