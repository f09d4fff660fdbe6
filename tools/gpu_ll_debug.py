import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import image_webp_b200 as Z
import oracle_lib as O
a = np.full((1, 5000, 4), 37, np.uint8)
a[0:, 2500:] = 200
a[0, 5] = 1
img = np.ascontiguousarray(a[:, :, :3])
ctx = Z.Context(0)
outs, t = ctx.encode_batch([img], Z.EncoderParams(), Z.ColorType.Rgb8, container=False)
ref = O.encode_lossless(img, "Rgb8")[1]
print("gpu", len(outs[0]), outs[0].hex())
print("ref", len(ref), ref.hex())
st = ctx.lossless_dump_stage(0, "LL_STATE", np.uint32)
print("state", st)
h = ctx.lossless_dump_stage(0, "LL_HEADER", np.uint32)
print("hdr words", [hex(x) for x in h[:24]])
hist = ctx.lossless_dump_stage(0, "LL_HIST", np.uint32).reshape(4, 280)
for c in range(4):
    print("hist", c, {int(i): int(v) for i, v in enumerate(hist[c]) if v})
codes = ctx.lossless_dump_stage(0, "LL_CODES", np.uint32).reshape(4, 280)
for c in range(4):
    print("codes", c, {int(i): (int(v) >> 16, int(v) & 0xffff) for i, v in enumerate(codes[c]) if v})
