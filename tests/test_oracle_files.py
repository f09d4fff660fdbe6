"""Whole-file checks of the CPU oracle: committed golden hashes (drift pin), the reference's own
acceptance tests for the lossy encoder (libwebp decodes it; PSNR / size thresholds,
tests/lossy_encoder_quality.rs), container bytes, error codes and edge cases."""
import hashlib
import io
import json
import os

import numpy as np
import pytest
from PIL import Image

import oracle_lib as O
import photo_inputs
from image_webp_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(255 * 255 / max(mse, 1e-9))


def decode(data):
    return np.array(Image.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_golden_hash(name):
    g = GOLD[name]
    img = (getattr(synth, g["kind"], None) or getattr(photo_inputs, g["kind"]))(*g["args"])
    assert hashlib.sha256(img.tobytes()).hexdigest() == g["input_sha256"], "input generator / photo fixture drifted"
    rc, data, _ = O.encode(img, g["quality"], g["method"])
    assert rc == 0 and len(data) == g["bytes"]
    assert hashlib.sha256(data).hexdigest() == g["sha256"]


def test_golden_webp_files_match_and_decode():
    for name in ("grad64_q75_m4", "photo99x87_s2_q75_m4"):
        data = open(os.path.join(HERE, "golden", name + ".webp"), "rb").read()
        assert hashlib.sha256(data).hexdigest() == GOLD[name]["sha256"]
        g = GOLD[name]
        img = getattr(synth, g["kind"])(*g["args"])
        dec = decode(data)
        assert dec.shape == img.shape and psnr(dec, img) > 30


def test_libwebp_can_decode_psnr_gradient():
    # tests/lossy_encoder_quality.rs:160-198 : 64x64 gradient, q75 -> libwebp decodes, PSNR > 15 dB
    img = synth.gradient(64, 64)
    rc, data, _ = O.encode(img, 75, 4)
    assert psnr(decode(data), img) > 15.0


@pytest.mark.parametrize("q", [50, 75, 90])
def test_size_and_quality_vs_libwebp(q):
    # tests/lossy_encoder_quality.rs:201-342 : size < 2.1x libwebp, PSNR >= 0.8x libwebp on 128x128 checker+gradient
    img = synth.checker_gradient(128, 128)
    rc, data, _ = O.encode(img, q, 4)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, "WEBP", quality=q, method=4)
    lw = buf.getvalue()
    assert len(data) < 2.1 * len(lw)
    assert psnr(decode(data), img) >= 0.8 * psnr(decode(lw), img)


def test_special_content_psnr():
    # tests/lossy_encoder_quality.rs:345-380 : solid / gradient / checkerboard > 20 dB, noise > 10 dB
    for img, thr in ((synth.solid(64, 64), 20), (synth.gradient(64, 64), 20), (synth.checker_gradient(64, 64, 8), 20), (synth.noise(64, 64), 10)):
        rc, data, _ = O.encode(img, 75, 4)
        assert psnr(decode(data), img) > thr


def test_container_layout():
    # src/encoder/api.rs:1224-1241, :1320-1329 and vp8.rs:315-330
    img = synth.photo_like(99, 87, 2)
    rc, webp, dump = O.encode(img, 75, 4, want_dump=True)
    rc2, vp8, _ = O.encode(img, 75, 4, container=False)
    assert webp[:4] == b"RIFF" and webp[8:12] == b"WEBP" and webp[12:16] == b"VP8 "
    n = int.from_bytes(webp[16:20], "little")
    assert n == len(vp8) and webp[20:20 + n] == vp8
    assert int.from_bytes(webp[4:8], "little") == len(webp) - 8
    assert len(webp) % 2 == 0
    tag = int.from_bytes(vp8[:3], "little")
    assert tag & 1 == 0 and (tag >> 4) & 1 == 1 and (tag >> 5) == dump["PART0"].size
    assert vp8[3:6] == bytes([0x9D, 0x01, 0x2A])
    assert int.from_bytes(vp8[6:8], "little") == 99 and int.from_bytes(vp8[8:10], "little") == 87
    assert len(vp8) == 10 + dump["PART0"].size + dump["PART1"].size


def test_error_codes():
    img = synth.photo_like(16, 16, 0)
    assert O.encode_raw(img.tobytes()[:-1], 16, 16, 75, 4)[0] == 2      # assert_eq! panic in the reference
    assert O.encode_raw(img.tobytes(), 70000, 1, 75, 4)[0] == 1         # InvalidDimensions
    assert O.encode_raw(img.tobytes(), 16, 16, 101, 4)[0] == 3          # quality panic


def test_method_clamp_and_segments_threshold():
    img = synth.photo_like(256, 256, 1)
    a = O.encode(img, 75, 6)[1]
    b = O.encode(img, 75, 9)[1]  # method.min(6), vp8.rs:1291
    assert a == b
    # 256 MBs -> segments enabled; 255 MBs (240x272) -> disabled (vp8.rs:2481)
    _, _, d1 = O.encode(img, 75, 4, want_dump=True)
    _, _, d2 = O.encode(synth.photo_like(240, 272, 1), 75, 4, want_dump=True)
    assert d1["SEG_ENABLED"][0] == 1 and "SEG_MAP" in d1
    assert d2["SEG_ENABLED"][0] == 0 and "SEG_MAP" not in d2


def test_pass1_pass2_records_are_consistent():
    _, _, d = O.encode(synth.photo_like(320, 272, 21), 75, 4, want_dump=True)
    p1, p2 = d["P1MB"], d["P2MB"]
    assert p1.shape == p2.shape == (20 * 17,)
    assert set(np.unique(p2["ymode"])) <= {0, 1, 2, 3, 4}
    assert (p2["bmodes"][p2["ymode"] != 4] == 0).all()
    assert (p2["levels"][p2["skip"] == 1] == 0).all()
    # Y2 block unused for B_PRED macroblocks
    assert (p2["levels"][p2["ymode"] == 4][:, 0, :] == 0).all()


def test_real_photographs_decode_and_have_photo_statistics():
    # the workload the reference benchmarks is a photograph (benches/profile_encode.rs:33): the oracle's output on
    # the reference's own test photos decodes with libwebp at a photo-grade PSNR, uses I4 nearly everywhere and
    # skips (almost) nothing -- unlike the synthetic generator (0.43 vs 1.1-2.3 token symbols per pixel)
    img = photo_inputs.survey_crop()
    rc, data, d = O.encode(img, 75, 4, want_dump=True)
    assert rc == 0 and psnr(decode(data), img) > 30
    nmb = d["P2MB"].size
    assert (d["P2MB"]["skip"] == 1).sum() < 0.02 * nmb
    assert d["TOK_TOKENS"].size > 1.5 * img.shape[0] * img.shape[1]   # > 1.5 symbols per pixel
    assert len(data) > 0.2 * img.shape[0] * img.shape[1]              # > 0.2 B/px


def test_batch_mt_entry_matches_single_encodes():
    b = photo_inputs.batch(3, 256, 192)
    outs = O.encode_batch_mt(b, 75, 4, threads=2)
    for i in range(3):
        assert outs[i] == O.encode(b[i], 75, 4)[1]
