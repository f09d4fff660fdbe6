"""Soak test of what cannot be race-checked on this pool (compute-sanitizer is closed there): the wavefront's progress
flags and the segment-parallel boolean coder under many schedules.  200+ encodes of one mixed-size batch -- different
persistent-grid sizes (warps per SM), pipeline depths, the three entry points, photo and synthetic content -- must all
produce ONE set of hashes, equal to the oracle's."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
import photo_inputs as PI
from image_webp_b200 import synth

pytestmark = pytest.mark.gpu


def _p(q, m):
    import image_webp_b200 as Z
    p = Z.EncoderParams.lossy(q)
    p.method = m
    return p


def _digest(outs):
    h = hashlib.sha256()
    for o in outs:
        h.update(hashlib.sha256(o).digest())
    return h.hexdigest()


def test_soak_one_hash_set_under_many_schedules():
    import image_webp_b200 as Z
    sizes = [(256, 256), (99, 87), (320, 272), (16, 16), (640, 360), (257, 255), (768, 64), (48, 512)]
    imgs = [synth.photo_like(w, h, 900 + i) for i, (w, h) in enumerate(sizes)]
    imgs += [PI.crop("3", 100 + 40 * i, 60 + 30 * i, 384, 256) for i in range(6)]
    imgs += [synth.noise(128, 128, 5), synth.solid(200, 120)]
    want = {}
    for q, m in ((75, 4), (50, 6), (90, 2)):
        want[(q, m)] = _digest([O.encode(im, q, m)[1] for im in imgs])
    runs = 0
    for warps in (0, 4, 8, 16, 24):
        for depth in (1, 3):
            ctx = Z.Context(0, persistent_warps_per_sm=warps, depth=depth)
            try:
                for rep in range(7):
                    for (q, m), ref in want.items():
                        if rep % 3 == 0:
                            outs, _ = ctx.encode_batch(imgs, _p(q, m))
                        elif rep % 3 == 1:
                            ctx.stage(imgs)
                            ctx.encode_resident(_p(q, m))
                            outs, _ = ctx.download()
                        else:
                            pend = [ctx.submit(imgs, _p(q, m)) for _ in range(depth)]
                            res = [p.result()[0] for p in pend]
                            assert all(_digest(r) == ref for r in res)
                            outs = res[-1]
                            runs += depth - 1
                        assert _digest(outs) == ref, "warps %d depth %d rep %d q%d m%d" % (warps, depth, rep, q, m)
                        runs += 1
            finally:
                ctx.close()
    assert runs >= 200
