#!/usr/bin/env python3
"""Where does the streaming path (zw_submit / zw_wait / zw_release, depth 3) lose time against the resident kernels?
Raw C ABI, no per-image Python work: wall per step, the per-batch device time the library reports, stage sums."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import image_webp_b200 as Z
from image_webp_b200 import synth, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 15
kind = sys.argv[4] if len(sys.argv) > 4 else "synthetic"
host = torch.empty((n, 512, 768, 3), dtype=torch.uint8, pin_memory=True)
if kind == "photo":
    import photo_inputs as PI
    PI.batch(n, out=host.numpy())
else:
    host.numpy()[...] = synth.batch_photo_like(n, 768, 512, 0)
L = _lib.load()
arr = (_lib.ZwImage * n)()
for i in range(n):
    arr[i] = _lib.ZwImage(host.numpy()[i].ctypes.data, 768 * 512 * 3, 768, 512, 2, 0)
ctx = Z.Context(0, depth=depth)
ctx.stage(list(host.numpy()))
p = Z.EncoderParams.lossy(75); p.method = 4
for _ in range(3):
    tr = ctx.encode_resident(p)
print("resident: device_total %.2f ms" % tr["device_total_ms"])
KEYS = ("yuv_ms", "analysis_ms", "pass1_ms", "chroma1_ms", "stats_ms", "chroma2_ms", "pass2_ms", "token_ms", "boolcode_ms", "assemble_ms")
def submit():
    t = C.c_int(-1)
    rc = L.zw_submit(ctx.h, arr, n, 75, 4, C.byref(t)); assert rc == 0, rc
    return t.value
def wait(tk):
    v = _lib.ZwBatchView(); tm = _lib.ZwTiming()
    rc = L.zw_wait(ctx.h, tk, 1, C.byref(v), C.byref(tm)); assert rc == 0, rc
    L.zw_release(ctx.h, tk)
    return tm.as_dict()
q = [submit() for _ in range(depth)]
for _ in range(depth): wait(q.pop(0)); q.append(submit())   # warm
while q: wait(q.pop(0))
torch.cuda.synchronize()
t0 = time.perf_counter()   # cold pipeline, like bench.py: the first batch's H2D is exposed, the last D2H too
tms = []
q = [submit() for _ in range(min(depth, steps))]
sub = len(q)
while q:
    tms.append(wait(q.pop(0)))
    if sub < steps:
        q.append(submit()); sub += 1
wall = time.perf_counter() - t0
dev = np.mean([t["device_total_ms"] for t in tms])
print("%s depth %d: wall %.2f ms/step; device_total per batch %.2f ms (resident %.2f); sum of stages %.2f" % (
    kind, depth, 1e3 * wall / steps, dev, tr["device_total_ms"], np.mean([sum(t[k] for k in KEYS) for t in tms])))
print("stages streaming: " + " ".join("%s %.2f" % (k[:-3], np.mean([t[k] for t in tms])) for k in KEYS))
print("stages resident : " + " ".join("%s %.2f" % (k[:-3], tr[k]) for k in KEYS))
print("h2d %.2f ms d2h %.2f ms" % (np.mean([t["h2d_ms"] for t in tms]), np.mean([t["d2h_ms"] for t in tms])))
