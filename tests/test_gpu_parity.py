"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, must be
byte-identical to the CPU oracle -- final .webp bytes and every intermediate stage."""
import hashlib
import io

import numpy as np
import pytest

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


def _params(q, m):
    import image_webp_b200 as Z
    p = Z.EncoderParams.lossy(q)
    p.method = m
    return p


CASES = [
    ("grad64_q75_m4", lambda: synth.gradient(64, 64), 75, 4),
    ("chk128_q75_m4", lambda: synth.checker_gradient(128, 128), 75, 4),
    ("solid64_q75_m4", lambda: synth.solid(64, 64), 75, 4),
    ("odd99x87_q75_m4", lambda: synth.photo_like(99, 87, 2), 75, 4),
    ("tiny1x1_q75_m4", lambda: synth.photo_like(1, 1, 3), 75, 4),
    ("tiny17x17_q50_m2", lambda: synth.photo_like(17, 17, 4), 50, 2),
    ("row_only_300x9_q75_m4", lambda: synth.photo_like(300, 9, 5), 75, 4),
    ("col_only_9x300_q75_m4", lambda: synth.photo_like(9, 300, 6), 75, 4),
    ("photo256_q50_m0", lambda: synth.photo_like(256, 256, 1), 50, 0),
    ("photo256_q75_m2", lambda: synth.photo_like(256, 256, 7), 75, 2),
    ("photo256_q75_m3", lambda: synth.photo_like(256, 256, 7), 75, 3),
    ("photo256_q75_m5", lambda: synth.photo_like(256, 256, 8), 75, 5),
    ("photo768_q75_m4", lambda: synth.photo_like(768, 512, 0), 75, 4),
    ("photo768_q75_m6", lambda: synth.photo_like(768, 512, 0), 75, 6),
    ("photo768_q90_m4", lambda: synth.photo_like(768, 512, 9), 90, 4),
    ("photo768_q20_m4", lambda: synth.photo_like(768, 512, 10), 20, 4),
    ("photo768_q100_m4", lambda: synth.photo_like(768, 512, 11), 100, 4),
    ("photo768_q0_m6", lambda: synth.photo_like(768, 512, 12), 0, 6),
    ("noise256_q90_m4", lambda: synth.noise(256, 256, 3), 90, 4),
    ("noise768_q90_m4", lambda: synth.noise(768, 512, 4), 90, 4),
    ("photo1080_q75_m6", lambda: synth.photo_like(1920, 1080, 13), 75, 6),
]


@pytest.mark.parametrize("name,gen,q,m", CASES, ids=[c[0] for c in CASES])
def test_single_image_byte_identical(ctx, name, gen, q, m):
    import image_webp_b200 as Z
    img = gen()
    ok, rep, gpu, ref = PU.check_image(ctx, img, q, m, Z.EncoderParams)
    assert ok, "%s: GPU output differs from the oracle (gpu %d B, oracle %d B)\n%s" % (name, len(gpu), len(ref), rep)


def test_all_stages_match_on_photo(ctx):
    img = synth.photo_like(320, 272, 21)
    rc, ref, dump = O.encode(img, 75, 4, want_dump=True)
    outs, _ = ctx.encode_batch([img], _params(75, 4))
    rep = PU.compare_stages(ctx, 0, dump, 20)
    assert not rep, "\n".join(rep)
    assert outs[0] == ref


def test_batch_mixed_sizes_identical_and_ordered(ctx):
    imgs = [synth.photo_like(w, h, s) for s, (w, h) in enumerate([(256, 256), (99, 87), (320, 272), (16, 16), (640, 360), (257, 255)] * 3)]
    outs, t = ctx.encode_batch(imgs, _params(75, 4))
    for i, im in enumerate(imgs):
        rc, ref, _ = O.encode(im, 75, 4)
        assert outs[i] == ref, "image %d of the batch differs" % i
    assert t["pixels"] == sum(im.shape[0] * im.shape[1] for im in imgs)


def test_batch_of_identical_config_images(ctx):
    imgs = list(synth.batch_photo_like(48, 256, 256, 100))
    outs, _ = ctx.encode_batch(imgs, _params(50, 0))
    for i in (0, 1, 17, 47):
        rc, ref, _ = O.encode(imgs[i], 50, 0)
        assert outs[i] == ref


def test_rgba_input_ignores_alpha(ctx):
    import image_webp_b200 as Z
    rgb = synth.photo_like(128, 96, 31)
    rgba = np.concatenate([rgb, np.full((96, 128, 1), 77, np.uint8)], axis=2)
    outs, _ = ctx.encode_batch([rgba], _params(75, 4), color=Z.ColorType.Rgba8, container=False)
    rc, ref, _ = O.encode(rgba, 75, 4, color="Rgba8", container=False)
    assert outs[0] == ref


@pytest.mark.parametrize("w,h,q,m", [(320, 272, 75, 4), (99, 87, 50, 2), (17, 33, 90, 6), (768, 512, 75, 4)])
def test_grey_input_l8_la8(ctx, w, h, q, m):
    # L8 / La8 go through convert_image_y (yuv.rs:806): Y = the grey sample, U = V = 127
    import image_webp_b200 as Z
    grey = synth.photo_like(w, h, 50 + w)[:, :, 1:2].copy()
    outs, _ = ctx.encode_batch([grey], _params(q, m), color=Z.ColorType.L8)
    rc, ref, dump = O.encode(grey, q, m, color="L8", want_dump=True)
    assert rc == 0
    if outs[0] != ref:
        pytest.fail("L8 differs:\n" + "\n".join(PU.compare_stages(ctx, 0, dump, (w + 15) // 16)[:4]))
    la = np.concatenate([grey, np.full((h, w, 1), 200, np.uint8)], axis=2)
    outs, _ = ctx.encode_batch([la], _params(q, m), color=Z.ColorType.La8, container=False)
    rc, ref2, _ = O.encode(la, q, m, color="La8", container=False)
    assert rc == 0 and outs[0] == ref2
    assert ref2 == ref[20:20 + len(ref2)]  # same VP8 payload as the L8 file: alpha is ignored by the VP8 path
    # WebPEncoder drop-in: L8 in the simple container, La8 as VP8X + ALPH + "VP8 " (api.rs:1330-1394)
    out = bytearray()
    enc = Z.WebPEncoder(out)
    enc.set_params(_params(q, m))
    enc.encode(grey.tobytes(), w, h, Z.ColorType.L8)
    assert bytes(out) == ref
    out = bytearray()
    enc = Z.WebPEncoder(out)
    enc.set_params(_params(q, m))
    enc.encode(la.tobytes(), w, h, Z.ColorType.La8)
    assert bytes(out) == O.webp_encode(la, "La8", use_lossy=True, quality=q, method=m)[1]


def test_decodes_with_libwebp_and_psnr(ctx):
    # the reference's own acceptance test: libwebp decodes our output, PSNR thresholds
    # (tests/lossy_encoder_quality.rs:160-198, :345-380)
    from PIL import Image

    def psnr(a, b):
        mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
        return 10 * np.log10(255 * 255 / max(mse, 1e-9))
    for img, thr in ((synth.gradient(64, 64), 15.0), (synth.solid(64, 64), 20.0), (synth.checker_gradient(128, 128), 20.0),
                     (synth.noise(64, 64), 10.0)):
        outs, _ = ctx.encode_batch([img], _params(75, 4))
        dec = np.array(Image.open(io.BytesIO(outs[0])).convert("RGB"))
        assert dec.shape == img.shape and psnr(dec, img) > thr


def test_error_codes(ctx):
    import image_webp_b200 as Z
    good = synth.photo_like(32, 32, 1)
    # wrong buffer size -> InvalidBufferSize (reference: assert_eq! panic, vp8.rs:1307)
    outs, _ = ctx.encode_batch([good, (bytes(100), 32, 32)], _params(75, 4), raise_errors=False)
    assert outs[0] is not None and outs[1] is None
    with pytest.raises(Z.InvalidBufferSize):
        ctx.encode_batch([(bytes(100), 32, 32)], _params(75, 4))
    with pytest.raises(Z.InvalidDimensions):
        ctx.encode_batch([(b"", 0, 5)], _params(75, 4))
    with pytest.raises(ValueError):
        ctx.encode_batch([good], _params(101, 4))


def test_webpencoder_dropin_appends(ctx):
    import image_webp_b200 as Z
    img = synth.photo_like(64, 48, 5)
    out = bytearray(b"prefix")
    enc = Z.WebPEncoder(out)
    p = Z.EncoderParams.lossy(75)
    p.method = 4
    enc.set_params(p)
    enc.encode(img.tobytes(), 64, 48, Z.ColorType.Rgb8)
    rc, ref, _ = O.encode(img, 75, 4)
    assert bytes(out) == b"prefix" + ref
    assert ref[:4] == b"RIFF" and ref[8:16] == b"WEBPVP8 "


def test_resident_split_api_matches_batch_api(ctx):
    imgs = [synth.photo_like(256, 256, 40 + i) for i in range(4)]
    ctx.stage(imgs)
    ctx.encode_resident(_params(75, 4))
    a, _ = ctx.download()
    t2 = ctx.encode_resident(_params(75, 4))  # re-encode the staged batch: idempotent
    b, _ = ctx.download()
    c, _ = ctx.encode_batch(imgs, _params(75, 4))
    assert a == b == c
    assert t2["kernel_launches"] >= 10


def test_batch_pipeline_matches_oracle_and_single_context(ctx):
    """The streaming entry point (several contexts / host threads on one GPU) returns, per batch and
    in order, the bytes the oracle produces."""
    import image_webp_b200 as Z
    batches = [[synth.photo_like(160 + 16 * k, 96 + 16 * j, 40 + 5 * k + j) for j in range(3)] for k in range(5)]
    with Z.BatchPipeline(0, depth=2) as pipe:
        got = pipe.encode_batches(batches, _params(75, 4))
    assert len(got) == len(batches)
    for b, outs in zip(batches, got):
        assert len(outs) == len(b)
        for img, o in zip(b, outs):
            rc, ref, _ = O.encode(img, 75, 4)
            assert rc == 0 and o == ref


def test_c_abi_refuses_alpha_in_the_simple_container(ctx):
    """zw_encode_webp_batch: lossy + alpha would need VP8X + ALPH (api.rs:1330-1394): per-image INVALID_PARAM, while
    the opaque images of the same batch are encoded."""
    import ctypes as C
    from image_webp_b200 import _lib
    L = _lib.load()
    rgb = synth.photo_like(64, 48, 3)
    rgba = np.concatenate([rgb, np.full((48, 64, 1), 9, np.uint8)], axis=2).copy()
    arr = (_lib.ZwImage * 2)()
    arr[0] = _lib.ZwImage(rgb.ctypes.data, rgb.size, 64, 48, 2, 0)
    arr[1] = _lib.ZwImage(rgba.ctypes.data, rgba.size, 64, 48, 3, 0)
    outs = (_lib.ZwOutput * 2)()
    rc = L.zw_encode_webp_batch(ctx.h, arr, 2, 75, 4, outs, None)
    assert rc == 0 and outs[0].status == 0 and outs[1].status == 3
    rc0, ref, _ = O.encode(rgb, 75, 4)
    assert C.string_at(outs[0].data, outs[0].len) == ref
    L.zw_free(outs[0].data)


def test_small_device_budget_splits_the_batch_into_chunks():
    """A context with a tiny device budget encodes the batch as several chunks (and, with lanes=2, two lanes
    per chunk): same bytes as the single-chunk context."""
    import image_webp_b200 as Z
    imgs = [synth.photo_like(128 + 16 * (i % 3), 96 + 16 * (i % 2), 70 + i) for i in range(12)]
    ref = []
    for im in imgs:
        rc, b, _ = O.encode(im, 75, 4)
        assert rc == 0
        ref.append(b)
    small = Z.Context(0, max_device_bytes=3 << 20)  # ~3 images per chunk
    try:
        outs, t = small.encode_batch(imgs, _params(75, 4))
        assert outs == ref
        assert t["kernel_launches"] > 17  # more than one chunk ran
    finally:
        small.close()
    two = Z.Context(0, lanes=2)
    try:
        outs, _ = two.encode_batch(imgs + imgs, _params(75, 4))
        assert outs == ref + ref
    finally:
        two.close()
