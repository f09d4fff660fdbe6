#!/usr/bin/env python3
"""Pins for the decoder oracle: the reference's own pixel-exact decode fixtures (tests/decode.rs:190-191).
Copies tests/images/gallery1/{1..5}.webp and records the SHA-256 of the RGB pixels of the reference PNGs
tests/reference/gallery1/*.png (default bilinear chroma upsampling) and tests/reference/gallery1_nofancy/*.png
(UpsamplingMethod::Simple), which the reference's decoder must reproduce with zero differing bytes.
Run in the build container (reads /root/reference); the GPU box only sees the committed outputs."""
import hashlib, json, os, shutil
import numpy as np
from PIL import Image
REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
os.makedirs(os.path.join(HERE, "decode"), exist_ok=True)
for i in range(1, 6):
    shutil.copyfile("%s/images/gallery1/%d.webp" % (REF, i), os.path.join(HERE, "decode", "%d.webp" % i))
    e = {}
    for key, d in (("fancy", "gallery1"), ("simple", "gallery1_nofancy")):
        im = np.asarray(Image.open("%s/reference/%s/%d.png" % (REF, d, i)).convert("RGB"))
        e[key] = {"sha256": hashlib.sha256(im.tobytes()).hexdigest(), "width": im.shape[1], "height": im.shape[0],
                  "sum": int(im.astype(np.uint64).sum())}
    out["gallery1/%d" % i] = e
# lossy + alpha files (VP8X: ALPH + 'VP8 ', NORMAL loop filter, level up to 63) and a 1x1 regression file: colour channels only
for name in ["gallery2/%d_webp_a" % i for i in range(1, 6)] + ["regression/dark"]:
    shutil.copyfile("%s/images/%s.webp" % (REF, name), os.path.join(HERE, "decode", name.replace("/", "_") + ".webp"))
    im = np.asarray(Image.open("%s/reference/%s.png" % (REF, name)).convert("RGBA"))[:, :, :3]
    im = np.ascontiguousarray(im)
    out[name] = {"fancy": {"sha256": hashlib.sha256(im.tobytes()).hexdigest(), "width": im.shape[1], "height": im.shape[0],
                           "sum": int(im.astype(np.uint64).sum())}}
json.dump(out, open(os.path.join(HERE, "decode_golden.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
