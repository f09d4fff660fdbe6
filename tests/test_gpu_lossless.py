"""GPU parity tests of the lossless (VP8L) path and of the complete WebPEncoder::encode mirror (run with -m gpu on a
B200): bytes identical to the lossless oracle at every stage, files that libwebp decodes back to the input pixels (the
reference's own acceptance test, api.rs:1405-1511), extended containers, the builder API."""
import io

import numpy as np
import pytest
from PIL import Image

import oracle_lib as O
import photo_inputs as PI
from image_webp_b200 import synth

pytestmark = pytest.mark.gpu
COLORS = ("Rgb8", "Rgba8", "L8", "La8")
MODE = {"Rgb8": "RGB", "Rgba8": "RGBA", "L8": "L", "La8": "LA"}


@pytest.fixture(scope="module")
def ctx():
    import image_webp_b200 as Z
    c = Z.Context(0)
    yield c
    c.close()


def _ct(color):
    import image_webp_b200 as Z
    return getattr(Z.ColorType, color)


def _view(rgba, color):
    return {"Rgba8": rgba, "Rgb8": rgba[:, :, :3], "La8": rgba[:, :, 1:3], "L8": rgba[:, :, 2]}[color].copy()


def _rgba(h, w, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if kind == "flat":
        a = np.full((h, w, 4), 37, np.uint8)
        a[h // 3:, w // 2:] = 200
        a[min(5, h - 1), min(5, w - 1)] = 1
        return a
    base = synth.photo_like(w, h, seed)
    alpha = (np.add.outer(np.arange(h), np.arange(w)) % 256).astype(np.uint8)
    alpha[h // 4: h // 2, w // 8: w // 2] = 0
    return np.dstack([base, alpha])


def _decode(data, mode):
    im = Image.open(io.BytesIO(data))
    im.load()
    return np.asarray(im.convert(mode))


SHAPES = [(1, 1, "noise"), (1, 5000, "flat"), (5000, 1, "noise"), (87, 99, "photo"), (64, 64, "noise"), (70, 300, "flat"),
          (512, 768, "photo"), (300, 5000, "flat"), (1023, 1025, "photo")]


@pytest.mark.parametrize("h,w,kind", SHAPES, ids=["%dx%d_%s" % (w, h, k) for h, w, k in SHAPES])
def test_lossless_bytes_and_roundtrip(ctx, h, w, kind):
    import image_webp_b200 as Z
    rgba = _rgba(h, w, 5, kind)
    for color in COLORS:
        img = _view(rgba, color)
        for pred in (True, False):
            p = Z.EncoderParams(use_predictor_transform=pred)
            outs, t = ctx.encode_batch([img], p, _ct(color))
            rc, ref = O.webp_encode(img, color, use_predictor=pred)
            assert rc == 0
            assert outs[0] == ref, (color, pred, len(outs[0]), len(ref))
            raw, _ = ctx.encode_batch([img], p, _ct(color), container=False)
            assert raw[0] == O.encode_lossless(img, color, use_predictor=pred)[1]
        got = _decode(outs[0], MODE[color])
        assert np.array_equal(got.reshape(img.shape), img)


def test_lossless_stage_dumps(ctx):
    import ctypes as C

    import image_webp_b200 as Z
    from test_lossless_host import H, FLAG_PRED
    img = _rgba(200, 333, 9, "photo")
    img[50:90] = img[50, 0]  # long runs, groups of 4097
    outs, _ = ctx.encode_batch([img], Z.EncoderParams(), Z.ColorType.Rgba8, container=False)
    n = img.shape[0] * img.shape[1]
    hist, codes = np.zeros(4 * 280, np.uint32), np.zeros(4 * 280, np.uint32)
    res, desc = np.zeros(n, np.uint32), np.zeros(n, np.uint16)
    out = np.zeros(n * 12 + 64, np.uint8)
    ln = H.hc_lossless(img.ctypes.data, 333, 200, 4, 3, FLAG_PRED, out.ctypes.data, out.size, hist.ctypes.data, codes.ctypes.data,
                       res.ctypes.data, desc.ctypes.data)
    assert np.array_equal(ctx.lossless_dump_stage(0, "LL_RESIDUAL", np.uint32), res)
    assert np.array_equal(ctx.lossless_dump_stage(0, "LL_TOKENS", np.uint16), desc)
    assert np.array_equal(ctx.lossless_dump_stage(0, "LL_HIST", np.uint32), hist)
    assert np.array_equal(ctx.lossless_dump_stage(0, "LL_CODES", np.uint32), codes)
    assert outs[0] == out[:ln].tobytes()


def test_lossless_mixed_batch(ctx):
    import image_webp_b200 as Z
    rng = np.random.default_rng(2)
    imgs = []
    for i in range(40):
        h, w = int(rng.integers(1, 200)), int(rng.integers(1, 300))
        imgs.append(_view(_rgba(h, w, i, ("photo", "noise", "flat")[i % 3]), "Rgb8"))
    outs, t = ctx.encode_batch(imgs, Z.EncoderParams(), Z.ColorType.Rgb8)
    for im, o in zip(imgs, outs):
        assert o == O.webp_encode(im, "Rgb8")[1]
    assert t["kernel_launches"] == 7 and t["pixels"] == sum(im.shape[0] * im.shape[1] for im in imgs)
    # several chunks (tiny device budget) give the same bytes
    c2 = Z.Context(0, max_device_bytes=1 << 20)
    outs2, t2 = c2.encode_batch(imgs, Z.EncoderParams(), Z.ColorType.Rgb8)
    c2.close()
    assert outs2 == outs and t2["kernel_launches"] > 7


def test_photo_batch_lossless(ctx):
    import image_webp_b200 as Z
    imgs = [PI.crop_origin(i) for i in range(12)]
    batch = PI.batch(12)
    outs, t = ctx.encode_batch(list(batch), Z.EncoderParams(), Z.ColorType.Rgb8)
    for i in range(12):
        assert outs[i] == O.webp_encode(batch[i], "Rgb8")[1]
        assert np.array_equal(_decode(outs[i], "RGB"), batch[i])


@pytest.mark.parametrize("color", ["Rgba8", "La8"])
def test_lossy_with_alpha_container(ctx, color):
    import image_webp_b200 as Z
    rgba = _rgba(272, 320, 4, "photo")
    img = _view(rgba, color)
    p = Z.EncoderParams.lossy(75)
    outs, _ = ctx.encode_batch([img, img[:100, :50].copy()], p, _ct(color))
    for im, o in zip([img, img[:100, :50].copy()], outs):
        rc, ref = O.webp_encode(im, color, use_lossy=True, quality=75, method=4)
        assert rc == 0 and o == ref
        got = _decode(o, MODE[color])
        assert np.array_equal(got[..., -1], im[..., -1])
    al, _ = ctx.encode_alpha_batch([img], _ct(color))
    assert al[0] == O.encode_alpha_lossless(img, color)[1]


def test_metadata_containers(ctx):
    import image_webp_b200 as Z
    rgb = _view(_rgba(60, 70, 8, "photo"), "Rgb8")
    rgba = _rgba(60, 70, 8, "photo")
    metas = [None, {"exif": b"0123456789"}, {"icc": b"i" * 9, "xmp": b"x" * 7}, {"icc": b"ab", "exif": b"c", "xmp": b"d" * 4}]
    for lossy in (False, True):
        p = Z.EncoderParams.lossy(60) if lossy else Z.EncoderParams()
        for color, img in (("Rgb8", rgb), ("Rgba8", rgba)):
            outs, _ = ctx.encode_batch([img] * len(metas), p, _ct(color), metadata=metas)
            for m, o in zip(metas, outs):
                m = m or {}
                rc, ref = O.webp_encode(img, color, use_lossy=lossy, quality=60, method=4, icc=m.get("icc", b""), exif=m.get("exif", b""),
                                        xmp=m.get("xmp", b""))
                assert rc == 0 and o == ref, (lossy, color, m)
                im = Image.open(io.BytesIO(o))
                im.load()
                if m.get("exif"):
                    assert im.info.get("exif") == m["exif"]
                if m.get("icc"):
                    assert im.info.get("icc_profile") == m["icc"]


def test_builder_api_and_errors(ctx):
    import image_webp_b200 as Z
    rgba = _rgba(33, 47, 1, "photo")
    out = Z.Encoder.new_rgba(rgba.tobytes(), 47, 33).quality(85.4).method(9).exif_metadata(b"EX").encode()
    assert out == O.webp_encode(rgba, "Rgba8", use_lossy=True, quality=85, method=6, exif=b"EX")[1]
    cfg = Z.EncoderConfig.new_lossless()
    assert cfg.is_lossless() and cfg.get_method() == 4 and cfg.get_quality() == 75.0
    assert cfg.encode_rgb(rgba[:, :, :3].tobytes(), 47, 33) == O.webp_encode(rgba[:, :, :3], "Rgb8")[1]
    buf = bytearray(b"keep")
    Z.Encoder.new_l8(rgba[:, :, 0].tobytes(), 47, 33).lossless(True).encode_into(buf)
    assert bytes(buf) == b"keep" + O.webp_encode(rgba[:, :, 0], "L8")[1]
    w = Z.WebPEncoder(bytearray())  # WebPEncoder::new defaults to lossless (api.rs:1256)
    w.encode(rgba.tobytes(), 47, 33, Z.ColorType.Rgba8)
    assert bytes(w.writer) == O.webp_encode(rgba, "Rgba8")[1]
    with pytest.raises(Z.InvalidBufferSize):
        Z.Encoder.new_rgb(b"\0" * 10, 47, 33).encode()
    with pytest.raises(Z.InvalidBufferSize):  # a larger buffer passes validate_buffer_size, then hits the size assert
        Z.Encoder.new_rgb(b"\0" * (47 * 33 * 3 + 1), 47, 33).lossless(True).encode()
    with pytest.raises(Z.InvalidDimensions):
        Z.Encoder.new_l8(b"", 0, 0).lossless(True).encode()
    wide = np.zeros((1, 16384), np.uint8)  # 16384 is legal for lossless only (api.rs:968 vs vp8.rs:3143)
    assert Z.Encoder.new_l8(wide.tobytes(), 16384, 1).lossless(True).encode() == O.webp_encode(wide, "L8")[1]
    with pytest.raises(Z.InvalidDimensions):
        Z.Encoder.new_l8(wide.tobytes(), 16384, 1).encode()


def test_length_limited_codes(ctx):
    # Fibonacci histograms: the unconstrained Huffman depth exceeds 15 (api.rs:225-262) -- the device takes the symbol
    # order of the limiting branch from a warp-wide rank sort; also the 7-bit limit of the code-length alphabet
    import image_webp_b200 as Z
    fib = [1, 1]
    while len(fib) < 27:
        fib.append(fib[-1] + fib[-2])
    vals = np.repeat(np.arange(27, dtype=np.uint8), fib)
    np.random.default_rng(5).shuffle(vals)
    w = 701
    h = len(vals) // w
    grey = vals[: w * h].reshape(h, w)
    rgb = np.dstack([grey, np.roll(grey, 3, axis=1), grey[::-1]])
    for img, color in ((grey, "L8"), (rgb, "Rgb8")):
        for pred in (False, True):
            outs, _ = ctx.encode_batch([img], Z.EncoderParams(use_predictor_transform=pred), _ct(color))
            assert outs[0] == O.webp_encode(img, color, use_predictor=pred)[1], (color, pred)
            assert np.array_equal(_decode(outs[0], MODE[color]).reshape(img.shape), img)


def test_raw_c_abi_mixed_colours_metadata_and_in_flight_batches(ctx):
    """zw_encode_batch through ctypes: one batch mixing all four colour types, metadata on some images only, lossy and
    lossless; called while a zw_submit batch of the same context is still in flight (the call waits for it; the ticket
    stays valid)."""
    import ctypes as C

    import image_webp_b200 as Z
    from image_webp_b200 import _lib
    L = _lib.load()
    rgba = _rgba(120, 161, 21, "photo")
    items = [("Rgba8", rgba), ("Rgb8", _view(rgba, "Rgb8")), ("La8", _view(rgba, "La8")), ("L8", _view(rgba, "L8")),
             ("Rgb8", _view(_rgba(33, 65, 3, "noise"), "Rgb8")), ("Rgba8", _rgba(1, 1, 4, "noise"))]
    metas = [None, {"exif": b"E" * 5}, {"icc": b"I" * 3, "xmp": b"X"}, None, None, {"xmp": b"xx"}]
    n = len(items)
    arr = (_lib.ZwImage * n)()
    keep = []
    for i, (color, im) in enumerate(items):
        buf = np.ascontiguousarray(im)
        keep.append(buf)
        arr[i] = _lib.ZwImage(buf.ctypes.data, buf.size, buf.shape[1], buf.shape[0], O.COLOR[color], 0)
    marr = (_lib.ZwMetadata * n)()
    for i, m in enumerate(metas):
        m = m or {}
        marr[i] = _lib.ZwMetadata(m.get("icc"), len(m.get("icc", b"")), m.get("exif"), len(m.get("exif", b"")), m.get("xmp"), len(m.get("xmp", b"")))
    lossy_p = Z.EncoderParams.lossy(70)
    pend = ctx.submit([_view(rgba, "Rgb8")] * 3, lossy_p)   # stays in flight across the lossless call below
    for lossy in (False, True):
        if lossy:  # the lossy batch call takes every pipeline slot: collect the ticket first
            files, _ = pend.result()
            assert files == [O.encode(_view(rgba, "Rgb8"), 70, 4)[1]] * 3
        zp = _lib.ZwParams(1, 1 if lossy else 0, 70, 4)
        outs = (_lib.ZwOutput * n)()
        rc = L.zw_encode_batch(ctx.h, arr, n, C.byref(zp), marr, outs, None)
        assert rc == 0
        for i, (color, im) in enumerate(items):
            m = metas[i] or {}
            ref = O.webp_encode(im, color, use_lossy=lossy, quality=70, method=4, icc=m.get("icc", b""), exif=m.get("exif", b""), xmp=m.get("xmp", b""))[1]
            assert outs[i].status == 0 and C.string_at(outs[i].data, outs[i].len) == ref, (lossy, i, color)
            L.zw_free(outs[i].data)
    # bad entries are reported per image, the rest of the batch is encoded
    bad = (_lib.ZwImage * 3)()
    bad[0] = arr[1]
    bad[1] = _lib.ZwImage(keep[1].ctypes.data, keep[1].size - 1, 161, 120, 2, 0)
    bad[2] = _lib.ZwImage(keep[3].ctypes.data, 0, 0, 0, 0, 0)
    outs = (_lib.ZwOutput * 3)()
    zp = L.zw_params_default()
    assert L.zw_encode_batch(ctx.h, bad, 3, C.byref(zp), None, outs, None) == 0
    assert [outs[i].status for i in range(3)] == [0, 2, 1]
    assert C.string_at(outs[0].data, outs[0].len) == O.webp_encode(items[1][1], "Rgb8")[1]
    L.zw_free(outs[0].data)


def test_lossless_large_image_and_repeat(ctx):
    # 16.8 Mpx in one image: 16384 tiles, carries / bit offsets across many warp-scan rounds; twice: same bytes
    import hashlib

    import image_webp_b200 as Z
    img = synth.photo_like(4096, 4096, 3, freq_scale=4.0)
    img[1000:1200] = img[1000, 0]       # 800 k identical pixels: run groups of 4097 across rows
    a, _ = ctx.encode_batch([img], Z.EncoderParams(), Z.ColorType.Rgb8)
    b, _ = ctx.encode_batch([img], Z.EncoderParams(), Z.ColorType.Rgb8)
    assert a == b
    ref = O.webp_encode(img, "Rgb8")[1]
    assert hashlib.sha256(a[0]).hexdigest() == hashlib.sha256(ref).hexdigest()
    assert np.array_equal(_decode(a[0], "RGB"), img)
