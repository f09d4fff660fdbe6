#!/usr/bin/env python3
"""Decoder workload for ncu: encode n photo crops (q75 m4), then zw_verify once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import image_webp_b200 as Z
import photo_inputs as PI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = Z.Context(0)
p = Z.EncoderParams.lossy(75); p.method = 4
pend = ctx.submit(list(PI.batch(n)), p)
info, ms = Z.verify_pending(pend)
pend.result()
print("verify n=%d: parse %.2f reconstruct %.2f filter %.2f colour %.2f ms" % ((n,) + tuple(ms)))
