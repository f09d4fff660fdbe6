"""GPU parity of the on-device VP8 decoder / batch verifier (zw_decode_batch, zw_verify) against the decoder oracle
(oracle/zw_dec_oracle.inc, pinned pixel-exact by the reference's decode fixtures): RGB with both upsampling methods,
filtered planes and macroblock modes (stage dumps), squared error against the source, error statuses; and the
reference's own fixtures straight against their golden hashes."""
import ctypes as C
import hashlib
import io
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import photo_inputs as PI
from image_webp_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "decode_golden.json")))


def _gfile(name):
    return open(os.path.join(HERE, "golden", "decode", name.replace("/", "_").replace("gallery1_", "") + ".webp"), "rb").read()


def _dump(ctx, index, stage, nbytes, dtype):
    buf = np.zeros(nbytes, np.uint8)
    ln = C.c_size_t(0)
    rc = ctx.lib.zw_decode_dump_stage(ctx.h, index, stage.encode(), buf.ctypes.data, buf.nbytes, C.byref(ln))
    assert rc == 0 and ln.value == nbytes
    return buf.view(dtype)


def _flags(mi):
    mi = mi.ravel()
    return (mi["luma_mode"].astype(np.uint32) | (mi["chroma_mode"].astype(np.uint32) << 3) | (mi["segment"].astype(np.uint32) << 5) |
            (mi["skipped"].astype(np.uint32) << 7) | (mi["non_zero_dct"].astype(np.uint32) << 8))


def test_reference_decode_fixtures_on_gpu():
    import image_webp_b200 as Z
    names = ["gallery1/%d" % i for i in range(1, 6)] + ["gallery2/%d_webp_a" % i for i in range(1, 6)] + ["regression/dark"]
    files = [_gfile(n) for n in names]
    for mode, ups in (("fancy", Z.UpsamplingMethod.Bilinear), ("simple", Z.UpsamplingMethod.Simple)):
        outs, info, _ = Z.decode_batch(files, ups)
        for n, px, inf in zip(names, outs, info):
            if mode not in GOLD[n]:
                continue
            g = GOLD[n][mode]
            assert inf["status"] == 0 and px.shape == (g["height"], g["width"], 3)
            assert hashlib.sha256(px.tobytes()).hexdigest() == g["sha256"], (n, mode)


def test_mixed_batch_matches_oracle_at_every_stage():
    import image_webp_b200 as Z
    from PIL import Image
    img = PI.crop("3", 256, 104, 384, 256)
    files, srcs = [], []
    for q, m in ((75, 4), (20, 4), (95, 6), (50, 0)):
        files.append(O.encode(img, q, m)[1]); srcs.append(img)
    for (w, h) in ((99, 87), (17, 17), (1, 1), (300, 9), (33, 250), (16, 16)):
        im = synth.photo_like(w, h, 7)
        files.append(O.encode(im, 60, 4, container=(w % 2 == 1))[1]); srcs.append(im)
    for q in (10, 50, 100):
        b = io.BytesIO()
        Image.fromarray(img).save(b, "WEBP", quality=q, method=4)
        files.append(b.getvalue()); srcs.append(img)
    files.append(O.encode(synth.noise(128, 128, 5), 90, 4)[1]); srcs.append(synth.noise(128, 128, 5))
    ctx = Z.Context(0)
    try:
        for ups in (Z.UpsamplingMethod.Bilinear, Z.UpsamplingMethod.Simple):
            # sources of different sizes cannot share one `sources` list with a fixed colour only if shapes differ -- they can: per-image w/h
            outs, info, ms = Z.decode_batch(files, ups, sources=srcs, ctx=ctx)
            assert ms[0] > 0
            for k, (f, s, px, inf) in enumerate(zip(files, srcs, outs, info)):
                rc, o = O.decode(f, ups == Z.UpsamplingMethod.Bilinear, ("rgb", "planes", "mbinfo"))
                assert rc == 0 and inf["status"] == 0
                assert np.array_equal(px, o["rgb"]), "rgb of file %d" % k
                assert inf["sse_rgb"] == int(((o["rgb"].astype(np.int64) - s.astype(np.int64)) ** 2).sum())
                assert inf["filter_level"] == o["hdr"]["filter_level"] and inf["filter_type"] == o["hdr"]["filter_type"]
                nmb = o["hdr"]["mbw"] * o["hdr"]["mbh"]
                planes = _dump(ctx, k, "DEC_PLANES", nmb * 384, np.uint8)
                op = np.concatenate([o["planes"]["y"].ravel(), o["planes"]["u"].ravel(), o["planes"]["v"].ravel()])
                assert np.array_equal(planes, op), "planes of file %d" % k
                mi = _dump(ctx, k, "DEC_MBINFO", nmb * 16, np.uint32).reshape(-1, 4)
                assert np.array_equal(mi[:, 0], _flags(o["mbinfo"])), "modes of file %d" % k
    finally:
        ctx.close()


def test_error_statuses_match_oracle():
    import image_webp_b200 as Z
    data = O.encode(synth.photo_like(64, 48, 1), 75, 4, container=False)[1]
    good = O.encode(synth.photo_like(48, 48, 2), 75, 4)[1]
    bad = [data[:3] + b"\x9d\x01\x2b" + data[6:], bytes([data[0] | 1]) + data[1:], data[:40], data[:len(data) // 2], data[:5],
           b"RIFF\x10\x00\x00\x00WEBPVP8L\x04\x00\x00\x00abcd", good]
    outs, info, _ = Z.decode_batch(bad, raise_errors=False)
    for f, px, inf in zip(bad, outs, info):
        rc = O.decode(f)[0]
        assert inf["status"] == rc, (inf, rc)
        assert (px is None) == (rc != 0)
    assert np.array_equal(outs[-1], O.decode(good, want=("rgb",))[1]["rgb"])
    with pytest.raises(Z.DecodingError):
        Z.decode_batch(bad[:1])


def test_webpdecoder_mirror():
    import image_webp_b200 as Z
    f = _gfile("gallery1/1")
    d = Z.WebPDecoder(f)
    assert d.dimensions() == (550, 368) and d.is_lossy() and not d.has_alpha() and not d.is_animated()
    assert hashlib.sha256(d.read_image().tobytes()).hexdigest() == GOLD["gallery1/1"]["fancy"]["sha256"]
    d.set_lossy_upsampling(Z.UpsamplingMethod.Simple)
    assert hashlib.sha256(d.read_image().tobytes()).hexdigest() == GOLD["gallery1/1"]["simple"]["sha256"]
    a = Z.WebPDecoder(_gfile("gallery2/1_webp_a"))
    assert a.has_alpha() and a.is_lossy() and a.output_buffer_size() == 400 * 301 * 4
    px, w, h = Z.decode_rgb(f)
    assert (w, h) == (550, 368)


def test_verify_in_place_after_encode():
    """zw_verify: the batch is decoded where the encoder left it in device memory and scored against the source pixels
    resident there -- same numbers as decoding the downloaded files with the oracle."""
    import image_webp_b200 as Z
    imgs = [PI.crop("3", 100 + 40 * i, 60 + 30 * i, 384, 256) for i in range(4)] + [synth.photo_like(99, 87, 3), synth.noise(48, 48, 1), synth.solid(40, 24)]
    ctx = Z.Context(0)
    try:
        for q, m in ((75, 4), (30, 6), (90, 0)):
            p = Z.EncoderParams.lossy(q)
            p.method = m
            pend = ctx.submit(imgs, p)
            info, ms = Z.verify_pending(pend)
            outs, _ = pend.result()
            for im, f, inf in zip(imgs, outs, info):
                rc, o = O.decode(f, True, ("rgb",))
                assert rc == 0 and inf["status"] == 0 and (inf["width"], inf["height"]) == (im.shape[1], im.shape[0])
                sse = int(((o["rgb"].astype(np.int64) - im.astype(np.int64)) ** 2).sum())
                assert inf["sse_rgb"] == sse
                if q == 75 and im.shape[0] >= 64:
                    assert inf["psnr_rgb"] > 24.0
    finally:
        ctx.close()


def test_verify_full_batch_psnr_floor():
    """Size-independent property at a BASELINE-sized batch: every one of 256 distinct photo crops encoded at q75 decodes on
    the device, and its PSNR against the source clears the reference's acceptance floor for this quality
    (tests/lossy_encoder_quality.rs thresholds are >= 28 dB luma-ish; RGB incl. chroma subsampling: > 24 dB)."""
    import image_webp_b200 as Z
    imgs = list(PI.batch(256))
    ctx = Z.Context(0)
    try:
        p = Z.EncoderParams.lossy(75)
        pend = ctx.submit(imgs, p)
        info, ms = Z.verify_pending(pend)
        pend.result()
        ps = np.array([i["psnr_rgb"] for i in info])
        assert all(i["status"] == 0 for i in info) and ps.min() > 24.0 and ps.max() < 60.0
    finally:
        ctx.close()
