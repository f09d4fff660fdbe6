"""Pins the DECODER oracle (oracle/zw_dec_oracle.inc, a restatement of src/decoder/{vp8,loop_filter,yuv,bit_reader}.rs):
(1) the reference's own pixel-exact decode fixtures -- tests/decode.rs:190-191 decodes tests/images/gallery1/1..5.webp and
    demands ZERO differing bytes against tests/reference/gallery1/*.png (bilinear upsampling) and gallery1_nofancy/*.png
    (UpsamplingMethod::Simple); the .webp files are committed under tests/golden/decode/, the PNG pixels as SHA-256
    (tests/golden/make_decode_golden.py);
(2) libwebp 1.6.0 (PIL), which the reference decoder states it matches ("dwebp's default conversion", tests/decode.rs:96):
    the gallery files all use the SIMPLE loop filter, so the NORMAL filter, hev thresholds and odd sizes are pinned on
    files produced by libwebp's encoder and by the encoder oracle."""
import hashlib
import io
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import photo_inputs as PI
from image_webp_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "decode_golden.json")))


def _file(i):
    return open(os.path.join(HERE, "golden", "decode", "%d.webp" % i), "rb").read()


@pytest.mark.parametrize("i", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("mode", ["fancy", "simple"])
def test_reference_decode_fixtures(i, mode):
    rc, out = O.decode(_file(i), fancy=(mode == "fancy"), want=("rgb",))
    g = GOLD["gallery1/%d" % i][mode]
    assert rc == 0
    assert out["rgb"].shape == (g["height"], g["width"], 3)
    assert hashlib.sha256(out["rgb"].tobytes()).hexdigest() == g["sha256"]


@pytest.mark.parametrize("name", ["gallery2/%d_webp_a" % i for i in range(1, 6)] + ["regression/dark"])
def test_reference_decode_fixtures_normal_filter(name):
    """tests/decode.rs:193,195-198: lossy + alpha files (VP8X, NORMAL loop filter, level 3..63) and a 1x1 image; colour
    channels of the reference PNGs (alpha is a VP8L plane, not part of this path)."""
    data = open(os.path.join(HERE, "golden", "decode", name.replace("/", "_") + ".webp"), "rb").read()
    rc, out = O.decode(data, want=("rgb",))
    g = GOLD[name]["fancy"]
    assert rc == 0 and out["hdr"]["filter_type"] == 0
    assert out["rgb"].shape == (g["height"], g["width"], 3)
    assert hashlib.sha256(out["rgb"].tobytes()).hexdigest() == g["sha256"]


def test_photo_fixture_pixels():
    """gallery1/3.png is committed in full (tests/golden/photos): compare pixels, not just the hash."""
    rc, out = O.decode(_file(3), want=("rgb",))
    assert rc == 0 and np.array_equal(out["rgb"], PI.photo("3"))


def _pil_decode(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def _libwebp_encode(img, **kw):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(img).save(b, "WEBP", **kw)
    return b.getvalue()


def test_matches_libwebp_on_normal_filter_files():
    img = PI.crop("3", 256, 104, 384, 256)
    files = [O.encode(img, q, m)[1] for q, m in ((75, 4), (20, 4), (95, 6), (50, 0))]
    files += [O.encode(synth.photo_like(w, h, 7), 60, 4)[1] for (w, h) in ((99, 87), (17, 17), (1, 1), (300, 9), (33, 250))]
    files += [O.encode(synth.noise(128, 128, 5), 90, 4)[1]]
    files += [_libwebp_encode(img, quality=q, method=m) for q in (10, 50, 100) for m in (0, 4)]
    files += [_libwebp_encode(PI.crop("5", 10, 20, 201, 133), quality=70, method=4)]
    kinds = set()
    for data in files:
        rc, out = O.decode(data, want=("rgb",))
        assert rc == 0
        kinds.add((out["hdr"]["filter_type"], out["hdr"]["filter_level"] >= 40, out["hdr"]["filter_level"] == 0))
        assert np.array_equal(out["rgb"], _pil_decode(data))
    assert (0, True, False) in kinds and (0, False, False) in kinds and (0, False, True) in kinds  # normal filter, hev 2 and < 2, no filter


def test_errors():
    data = O.encode(synth.photo_like(64, 48, 1), 75, 4, container=False)[1]
    assert O.decode(data)[0] == 0
    assert O.decode(data[:2])[0] == 5                                  # truncated tag
    assert O.decode(data[:3] + b"\x9d\x01\x2b" + data[6:])[0] == 3    # bad start code
    assert O.decode(bytes([data[0] | 1]) + data[1:])[0] == 2          # inter frame
    assert O.decode(data[:40])[0] in (1, 5)                           # first partition cut short / bitstream error
    assert O.decode(b"RIFF\x10\x00\x00\x00WEBPVP8L\x04\x00\x00\x00abcd")[0] == 6  # no 'VP8 ' chunk


def test_unfiltered_planes_and_modes_are_consistent():
    img = PI.crop("4", 100, 60, 160, 96)
    data = O.encode(img, 30, 4)[1]
    rc, out = O.decode(data, want=("planes", "planes_unfiltered", "mbinfo"))
    assert rc == 0 and out["hdr"]["filter_level"] > 0
    assert not np.array_equal(out["planes"]["y"], out["planes_unfiltered"]["y"])
    assert out["mbinfo"].shape == (6, 10) and set(np.unique(out["mbinfo"]["luma_mode"])) <= {0, 1, 2, 3, 4}
