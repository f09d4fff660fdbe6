"""Pins the lossless (VP8L) half of the oracle the way the reference pins its own encoder: libwebp must decode every
file to exactly the input pixels (src/encoder/api.rs:1405-1511: write_webp, write_webp_exif, roundtrip_libwebp with and
without the predictor transform, with ICC / EXIF / XMP chunks).  libwebp 1.6.0 comes with PIL."""
import io
import struct

import numpy as np
import pytest
from PIL import Image

import oracle_lib as O
from image_webp_b200 import synth


def _decode(data, mode):
    im = Image.open(io.BytesIO(data))
    im.load()
    return np.asarray(im.convert(mode)), im


def _chunks(data):
    assert data[:4] == b"RIFF" and data[8:12] == b"WEBP"
    assert struct.unpack("<I", data[4:8])[0] == len(data) - 8
    pos, out = 12, []
    while pos < len(data):
        name, n = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        out.append((name, data[pos + 8:pos + 8 + n]))
        pos += 8 + n + (n & 1)
    assert pos == len(data)
    return out


def _noise(h, w, c, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, c) if c > 1 else (h, w), dtype=np.uint8)


@pytest.mark.parametrize("pred", [True, False])
@pytest.mark.parametrize("color,c,mode", [("Rgb8", 3, "RGB"), ("Rgba8", 4, "RGBA"), ("L8", 1, "L"), ("La8", 2, "LA")])
def test_roundtrip_libwebp_noise(color, c, mode, pred):
    # api.rs:1456-1511 (random 256x256, Rgb8 and Rgba8, both predictor settings); grey types added
    img = _noise(256, 256, c, 7)
    rc, data = O.webp_encode(img, color, use_predictor=pred)
    assert rc == 0
    names = [n for n, _ in _chunks(data)]
    assert names == [b"VP8L"]
    got, _ = _decode(data, mode)
    assert np.array_equal(got.reshape(img.shape), img)


@pytest.mark.parametrize("pred", [True, False])
@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (9, 1), (87, 99), (64, 5000), (300, 300)])
def test_roundtrip_shapes_and_runs(shape, pred):
    h, w = shape
    # photo-like content, flat areas (long runs incl. > 4096 and run groups of 4097), and a run across a row end
    img = synth.photo_like(w, h, 3) if min(h, w) >= 16 else _noise(h, w, 3, 1)
    img = img.copy()
    if h >= 64:
        img[10:40] = 77           # flat rows: residual runs longer than 4096 pixels
        img[50, w // 2:] = 5
        img[51, : w // 3] = 5
    rc, data = O.webp_encode(img, "Rgb8", use_predictor=pred)
    assert rc == 0
    got, _ = _decode(data, "RGB")
    assert np.array_equal(got, img)
    rgba = np.dstack([img, np.full((h, w), 200, np.uint8)])
    rgba[h // 2:, :, 3] = 13
    rc, data = O.webp_encode(rgba, "Rgba8", use_predictor=pred)
    assert rc == 0
    got, _ = _decode(data, "RGBA")
    assert np.array_equal(got, rgba)


def test_flat_image_single_symbol_trees():
    img = np.full((40, 33, 3), 9, np.uint8)
    for pred in (True, False):
        rc, data = O.webp_encode(img, "Rgb8", use_predictor=pred)
        assert rc == 0
        got, _ = _decode(data, "RGB")
        assert np.array_equal(got, img)


def test_metadata_chunks_and_order():
    # api.rs:1330-1394: VP8X, ICCP, (ALPH), frame, EXIF, XMP; flags; canvas size
    img = _noise(31, 45, 4, 3)
    rc, data = O.webp_encode(img, "Rgba8", icc=b"i" * 9, exif=b"e" * 10, xmp=b"x" * 7)
    assert rc == 0
    ch = _chunks(data)
    assert [n for n, _ in ch] == [b"VP8X", b"ICCP", b"VP8L", b"EXIF", b"XMP "]
    vp8x = ch[0][1]
    assert vp8x[0] == (1 << 2) | (1 << 3) | (1 << 4) | (1 << 5) and vp8x[1:4] == b"\0\0\0"
    assert int.from_bytes(vp8x[4:7], "little") == 44 and int.from_bytes(vp8x[7:10], "little") == 30
    assert ch[1][1] == b"i" * 9 and ch[3][1] == b"e" * 10 and ch[4][1] == b"x" * 7
    got, im = _decode(data, "RGBA")
    assert np.array_equal(got, img)
    assert im.info.get("exif") == b"e" * 10 and im.info.get("icc_profile") == b"i" * 9
    # opaque + EXIF only (write_webp_exif, api.rs:1423)
    rgb = _noise(20, 20, 3, 4)
    rc, data = O.webp_encode(rgb, "Rgb8", exif=b"0123456789")
    assert [n for n, _ in _chunks(data)] == [b"VP8X", b"VP8L", b"EXIF"]
    assert _chunks(data)[0][1][0] == 1 << 3
    got, _ = _decode(data, "RGB")
    assert np.array_equal(got, rgb)


@pytest.mark.parametrize("color,c,mode", [("Rgba8", 4, "RGBA"), ("La8", 2, "LA")])
def test_lossy_with_alpha(color, c, mode):
    # api.rs:1296, :1352-1360: VP8X + ALPH (lossless, implicit dimensions, predictor on) + "VP8 "
    base = synth.photo_like(99, 87, 5)
    alpha = (np.add.outer(np.arange(87), np.arange(99)) % 256).astype(np.uint8)
    alpha[20:40, 10:60] = 0
    img = np.dstack([base, alpha]) if c == 4 else np.dstack([base[:, :, 1], alpha])
    rc, data = O.webp_encode(img, color, use_lossy=True, quality=75, method=4)
    assert rc == 0
    ch = _chunks(data)
    assert [n for n, _ in ch] == [b"VP8X", b"ALPH", b"VP8 "]
    assert ch[0][1][0] == 1 << 4
    assert ch[1][1][0] == 1  # no preprocessing, no filter, lossless compression
    rc2, alph = O.encode_alpha_lossless(img, color)
    assert rc2 == 0 and alph == ch[1][1]
    rc3, vp8, _ = O.encode(img, 75, 4, color=color, container=False)
    assert rc3 == 0 and vp8 == ch[2][1]
    got, _ = _decode(data, mode)
    assert np.array_equal(got[..., -1], alpha)  # alpha is lossless
    ref = img[..., :3] if c == 4 else img[..., 0]
    err = got[..., :-1].astype(np.float64).reshape(87, 99, -1) - ref.astype(np.float64).reshape(87, 99, -1)
    assert 10 * np.log10(255 ** 2 / (err ** 2).mean()) > 30


def test_errors():
    assert O.webp_encode(None, "Rgb8", raw=b"\0" * 11, w=2, h=2)[0] == 2
    assert O.webp_encode(None, "Rgb8", raw=b"", w=0, h=0)[0] == 1
    rc, _ = O.encode_lossless(np.zeros((1, 16384, 1), np.uint8)[:, :, 0], "L8")
    assert rc == 0  # 16384 is allowed for lossless (api.rs:970), unlike the lossy path


def test_huffman_properties():
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = 280 if trial % 2 else 256
        f = np.zeros(n, np.uint32)
        k = int(rng.integers(2, n + 1))
        idx = rng.choice(n, k, replace=False)
        if trial % 3 == 0:
            f[idx] = rng.integers(1, 4, k)
        elif trial % 3 == 1:
            f[idx] = (rng.pareto(0.4, k) * 3 + 1).clip(1, 2 ** 28).astype(np.uint32)  # heavy tail: exercises the 15-bit limit
        else:
            f[idx] = rng.integers(1, 100000, k)
        ok, lengths, codes = O.build_huffman(f, 15)
        assert ok
        assert ((lengths > 0) == (f > 0)).all() and lengths.max() <= 15
        assert abs(sum(2.0 ** -int(l) for l in lengths if l) - 1.0) < 1e-12  # complete prefix code
        # canonical: codes (bit-reversed back) increase with (length, index)
        order = sorted((int(l), i) for i, l in enumerate(lengths) if l)
        vals = [int(format(int(codes[i]), "0%db" % l)[::-1], 2) << (15 - l) for l, i in order]
        assert vals == sorted(vals) and len(set(vals)) == len(vals)
    # Fibonacci frequencies force depth > 15 -> the limiting branch
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    f = np.zeros(256, np.uint32)
    f[:30] = fib
    ok, lengths, _ = O.build_huffman(f, 15)
    assert ok and lengths[:30].max() == 15 and abs(sum(2.0 ** -int(l) for l in lengths if l) - 1.0) < 1e-12
    assert not O.build_huffman(np.array([0, 5, 0, 0], np.uint32), 7)[0]


def test_length_limited_codes_decode():
    # Fibonacci histogram of grey values: the unconstrained Huffman depth exceeds 15 (api.rs:225-262)
    fib = [1, 1]
    while len(fib) < 27:
        fib.append(fib[-1] + fib[-2])
    vals = np.repeat(np.arange(27, dtype=np.uint8), fib)
    np.random.default_rng(5).shuffle(vals)
    w = 701
    h = len(vals) // w
    img = vals[: w * h].reshape(h, w)
    for pred in (False, True):
        rc, data = O.webp_encode(img, "L8", use_predictor=pred)
        assert rc == 0
        got, _ = _decode(data, "L")
        assert np.array_equal(got, img)
