#!/usr/bin/env python3
"""e2e pipeline throughput with minimal Python work per step: prebuilt zw_image array, caller-provided output buffers."""
import os, sys, time, threading, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import image_webp_b200 as Z
from image_webp_b200 import synth, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
host = torch.empty((n, 512, 768, 3), dtype=torch.uint8, pin_memory=True)
host.numpy()[...] = synth.batch_photo_like(n, 768, 512, 0)
L = _lib.load()
arr = (_lib.ZwImage * n)()
for i in range(n):
    arr[i] = _lib.ZwImage(host.numpy()[i].ctypes.data, 768 * 512 * 3, 768, 512, 2, 0)
CAP = 96 * 1024
ctxs = [Z.Context(0) for _ in range(depth)]
obufs = [np.zeros((n, CAP), np.uint8) for _ in range(depth)]
outs = []
for k in range(depth):
    o = (_lib.ZwOutput * n)()
    for i in range(n):
        o[i].data = obufs[k][i].ctypes.data; o[i].cap = CAP
    outs.append(o)
def run(k):
    t = _lib.ZwTiming()
    rc = L.zw_encode_webp_batch(ctxs[k].h, arr, n, 75, 4, outs[k], C.byref(t))
    assert rc == 0
for k in range(depth): run(k)
def worker(k):
    for s in range(k, steps, depth): run(k)
t0 = time.perf_counter()
th = [threading.Thread(target=worker, args=(k,)) for k in range(depth)]
for t in th: t.start()
for t in th: t.join()
wall = time.perf_counter() - t0
print("n%d depth%d steps%d (raw C ABI, caller buffers): %.1f ms/step -> e2e %.0f MPix/s" % (n, depth, steps, 1e3 * wall / steps, steps * n * 768 * 512 / wall / 1e6))
