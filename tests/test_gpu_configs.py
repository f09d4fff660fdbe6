"""GPU parity at the BASELINE.json configuration shapes (run with -m gpu).

Full-size oracle comparisons where the oracle finishes in seconds (one 4096x4096 image, a few
1080p images, samples of the large batches), plus size-independent properties for whole batches:
every output decodes with libwebp, batch outputs are independent of batch composition / order
(encode(batch)[i] == encode([img_i])), re-encoding is idempotent."""
import hashlib
import io

import numpy as np
import pytest
from PIL import Image

import oracle_lib as O
import parity_util as PU
from image_webp_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import image_webp_b200 as Z
    c = Z.Context(0)
    yield c
    c.close()


def _p(q, m):
    import image_webp_b200 as Z
    p = Z.EncoderParams.lossy(q)
    p.method = m
    return p


def test_config3_single_4096_q90_m4(ctx):
    # one 4096x4096 image: 65536 macroblocks, statistics counters overflow -> ordered halving (Q9)
    img = synth.photo_like(4096, 4096, 3, freq_scale=4.0)
    rc, ref, dump = O.encode(img, 90, 4, want_dump=True)
    assert rc == 0
    assert dump["PART0"].size < (1 << 19)
    assert (dump["STATS"] >> 16).max() >= 32767  # the halving path is exercised
    outs, t = ctx.encode_batch([img], _p(90, 4))
    if outs[0] != ref:
        rep = PU.compare_stages(ctx, 0, dump, 256)
        pytest.fail("4096x4096 differs:\n" + "\n".join(rep[:4]))


def test_config3_noisy_4096_statistics_halving(ctx):
    # busier content: many slots halve several times
    rng = np.random.default_rng(5)
    base = synth.photo_like(2048, 2048, 4)
    img = np.clip(base.astype(np.int16) + rng.integers(-40, 41, base.shape), 0, 255).astype(np.uint8)
    rc, ref, dump = O.encode(img, 75, 2, want_dump=True)
    assert (dump["STATS"] >> 16).max() >= 32767
    outs, _ = ctx.encode_batch([img], _p(75, 2))
    if outs[0] != ref:
        rep = PU.compare_stages(ctx, 0, dump, 128)
        pytest.fail("noisy 2048x2048 differs:\n" + "\n".join(rep[:4]))


def test_config4_1080p_m6_batch(ctx):
    imgs = [synth.photo_like(1920, 1080, 50 + i) for i in range(4)]
    outs, _ = ctx.encode_batch(imgs, _p(75, 6))
    for i in (0, 3):
        rc, ref, _ = O.encode(imgs[i], 75, 6)
        assert outs[i] == ref
    for o in outs:
        assert Image.open(io.BytesIO(o)).size == (1920, 1080)


def test_config5_thumbnails_m0_batch(ctx):
    imgs = list(synth.batch_photo_like(512, 256, 256, 200))
    outs, _ = ctx.encode_batch(imgs, _p(50, 0))
    for i in (0, 63, 64, 200, 511):
        rc, ref, _ = O.encode(imgs[i], 50, 0)
        assert outs[i] == ref
    assert all(o[:4] == b"RIFF" for o in outs)


def test_config2_batch_properties(ctx):
    imgs = list(synth.batch_photo_like(96, 768, 512, 300))
    outs, _ = ctx.encode_batch(imgs, _p(75, 4))
    # oracle on a sample
    for i in (0, 31, 95):
        rc, ref, _ = O.encode(imgs[i], 75, 4)
        assert outs[i] == ref
    # independence of batch composition and order
    perm = [95, 3, 40, 17, 64]
    sub, _ = ctx.encode_batch([imgs[i] for i in perm], _p(75, 4))
    assert [hashlib.sha256(x).digest() for x in sub] == [hashlib.sha256(outs[i]).digest() for i in perm]
    # idempotence
    again, _ = ctx.encode_batch(imgs, _p(75, 4))
    assert again == outs
    # every file decodes at the right size with a sane PSNR
    for i in (5, 50):
        dec = np.array(Image.open(io.BytesIO(outs[i])).convert("RGB"))
        mse = np.mean((dec.astype(np.float64) - imgs[i].astype(np.float64)) ** 2)
        assert dec.shape == imgs[i].shape and 10 * np.log10(255 * 255 / mse) > 28
