#!/usr/bin/env python3
"""Soak of the lossless path: many mixed batches (sizes, colour types, predictor on/off, chunk splits), every file compared
with the lossless oracle; then the same batch 30 times: one hash set."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import image_webp_b200 as Z
import oracle_lib as O
from image_webp_b200 import synth
rng = np.random.default_rng(7)
ctx = Z.Context(0)
CT = {"Rgb8": Z.ColorType.Rgb8, "Rgba8": Z.ColorType.Rgba8, "L8": Z.ColorType.L8, "La8": Z.ColorType.La8}
n_files = 0
for rnd in range(24):
    color = ("Rgb8", "Rgba8", "L8", "La8")[rnd % 4]
    pred = bool(rnd & 4) or rnd < 4
    imgs = []
    for i in range(int(rng.integers(1, 40))):
        h, w = int(rng.integers(1, 400)), int(rng.integers(1, 500))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            rgba = np.dstack([synth.photo_like(w, h, rnd * 100 + i), (np.add.outer(np.arange(h), np.arange(w)) % 256).astype(np.uint8)])
        elif kind == 1:
            rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        else:
            rgba = np.full((h, w, 4), int(rng.integers(0, 256)), np.uint8); rgba[h // 2:, : w // 3] = 9
        imgs.append({"Rgba8": rgba, "Rgb8": rgba[:, :, :3], "La8": rgba[:, :, 1:3], "L8": rgba[:, :, 2]}[color].copy())
    os.environ["ZW_LL_SPLIT"] = str(1 + rnd % 5)
    outs, _ = ctx.encode_batch(imgs, Z.EncoderParams(use_predictor_transform=pred), CT[color])
    for im, o in zip(imgs, outs):
        assert o == O.webp_encode(im, color, use_predictor=pred)[1], (rnd, color, pred, im.shape)
    n_files += len(imgs)
os.environ.pop("ZW_LL_SPLIT", None)
batch = [synth.photo_like(768, 512, i) for i in range(96)]
hs = set()
for rep in range(30):
    outs, _ = ctx.encode_batch(batch, Z.EncoderParams(), Z.ColorType.Rgb8)
    hs.add(hashlib.sha256(b"".join(hashlib.sha256(o).digest() for o in outs)).hexdigest())
assert len(hs) == 1, hs
print("lossless soak ok: %d files identical to the oracle, 30 repeats of a 96-image batch -> one hash" % n_files)
