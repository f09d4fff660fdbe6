#!/usr/bin/env python3
"""The lossless workload once, for ncu: n photo crops (768x512 RGB) through zw_encode_batch (lossless default)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import image_webp_b200 as Z
import photo_inputs as PI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
imgs = list(PI.batch(n))
ctx = Z.Context(0)
outs, t = ctx.encode_batch(imgs, Z.EncoderParams(), Z.ColorType.Rgb8)
print("lossless n=%d: device %.2f ms, %.3f B/px" % (n, t["device_total_ms"], sum(len(o) for o in outs) / (n * 768 * 512)))
