"""GPU parity on REAL photographs (run with -m gpu): crops of the reference's own test photos
(tests/golden/photos = /root/reference/tests/reference/gallery1/{3,4,5}.png).  Photographs carry
2-5x the token symbols per pixel of the synthetic generator, (almost) no skipped macroblocks and
I4 nearly everywhere: the tokeniser, boolean coder, trellis and statistics run at rates the
synthetic cases never reach.  Also covers method 1 and method 9 (clamped to 6, vp8.rs:1291)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import parity_util as PU
import photo_inputs as PI

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")))


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


CASES = [
    ("real3_q75_m4", lambda: PI.survey_crop(), 75, 4),
    ("real3_q75_m6", lambda: PI.survey_crop(), 75, 6),
    ("real4_q90_m4", lambda: PI.crop("4", 100, 200), 90, 4),
    ("real5_q50_m0", lambda: PI.crop("5", 0, 0), 50, 0),
    ("real5_q75_m1", lambda: PI.crop("5", 200, 240), 75, 1),
    ("real4_q75_m9_clamped", lambda: PI.crop("4", 17, 33), 75, 9),
    ("real3_q75_m2", lambda: PI.crop("3", 512, 208), 75, 2),
    ("real3_q75_m3", lambda: PI.crop("3", 0, 0), 75, 3),
    ("real5_q75_m5", lambda: PI.crop("5", 256, 0), 75, 5),
    ("real4_q100_m4", lambda: PI.crop("4", 256, 260), 100, 4),
    ("real3_q20_m4", lambda: PI.crop("3", 300, 50), 20, 4),
    ("real3_full1280x720_q75_m4", lambda: PI.photo("3"), 75, 4),
    ("real4_full1024x772_q75_m4", lambda: PI.photo("4"), 75, 4),
    ("real5_full1024x752_q90_m6", lambda: PI.photo("5"), 90, 6),
    ("real3_odd333x211_q75_m4", lambda: PI.crop("3", 400, 300, 333, 211), 75, 4),
]


@pytest.mark.parametrize("name,gen,q,m", CASES, ids=[c[0] for c in CASES])
def test_photo_byte_identical(ctx, name, gen, q, m):
    import image_webp_b200 as Z
    img = gen()
    ok, rep, gpu, ref = PU.check_image(ctx, img, q, m, Z.EncoderParams)
    assert ok, "%s: GPU output differs from the oracle (gpu %d B, oracle %d B)\n%s" % (name, len(gpu), len(ref), rep)


def test_photo_matches_committed_golden_hashes(ctx):
    # the same files the CPU suite pins for the oracle (tests/golden/oracle_golden.json)
    for name, g in GOLD.items():
        if g["kind"] not in ("crop", "photo"):
            continue
        img = getattr(PI, g["kind"])(*g["args"])
        outs, _ = ctx.encode_batch([img], _p(g["quality"], g["method"]))
        assert hashlib.sha256(outs[0]).hexdigest() == g["sha256"], name


def test_photo_all_stages(ctx):
    img = PI.survey_crop()
    rc, ref, dump = O.encode(img, 75, 4, want_dump=True)
    outs, _ = ctx.encode_batch([img], _p(75, 4))
    rep = PU.compare_stages(ctx, 0, dump, 48)
    assert not rep, "\n".join(rep[:4])
    assert outs[0] == ref


def test_photo_batch_every_image_checked(ctx):
    # 96 distinct crops in one batch, EVERY output compared with the multi-threaded oracle
    b = PI.batch(96)
    outs, t = ctx.encode_batch(list(b), _p(75, 4))
    ref = O.encode_batch_mt(b, 75, 4)
    bad = [i for i in range(len(ref)) if outs[i] != ref[i]]
    assert not bad, "images %s of the photo batch differ" % bad[:8]
    assert t["symbols"] > 1.0 * t["pixels"]  # photo token rate (synthetic: ~0.43 per pixel)
