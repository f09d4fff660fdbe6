"""CPU check of the on-device decoder: image_webp_b200/csrc/zw_dec.cuh -- the SAME source the kernels run -- is compiled with
g++ (tests/hostcheck) and executed lane by lane on the host, then compared with the decoder oracle: filtered planes,
macroblock modes / flags, RGB (bilinear and nearest upsampling) and the squared error against the source.  Catches
logic slips before any GPU time is spent; the `-m gpu` twin is tests/test_gpu_decoder.py."""
import ctypes as C
import io
import os

import numpy as np
import pytest

import oracle_lib as O
import photo_inputs as PI
from image_webp_b200 import synth
from test_device_prims_host import H

HERE = os.path.dirname(os.path.abspath(__file__))
H.hc_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]


def strip_container(data):
    """The 'VP8 ' chunk of a RIFF/WEBP file (simple or VP8X), or the bytes themselves."""
    if data[:4] != b"RIFF":
        return data
    pos = 12
    while pos + 8 <= len(data):
        sz = int.from_bytes(data[pos + 4:pos + 8], "little")
        if data[pos:pos + 4] == b"VP8 ":
            return data[pos + 8:pos + 8 + sz]
        pos += 8 + sz + (sz & 1)
    raise ValueError("no VP8 chunk")


def hc_decode(data, fancy=True, src=None):
    vp8 = strip_container(data)
    w, h = (vp8[6] | vp8[7] << 8) & 0x3FFF, (vp8[8] | vp8[9] << 8) & 0x3FFF
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    planes = np.zeros(mbw * mbh * 384, np.uint8)
    mbinfo = np.zeros((mbw * mbh, 4), np.uint32)
    rgb = np.zeros((h, w, 3), np.uint8)
    st = (C.c_uint32 * 12)()
    sse = C.c_uint64(0)
    rc = H.hc_decode(vp8, len(vp8), w, h, 1 if fancy else 0, planes.ctypes.data, mbinfo.ctypes.data, rgb.ctypes.data,
                     src.ctypes.data if src is not None else None, src.shape[2] if src is not None else 0, st, C.byref(sse))
    return rc, planes, mbinfo, rgb, list(st), sse.value


def oracle_flags(mi):
    mi = mi.ravel()
    return (mi["luma_mode"].astype(np.uint32) | (mi["chroma_mode"].astype(np.uint32) << 3) | (mi["segment"].astype(np.uint32) << 5) |
            (mi["skipped"].astype(np.uint32) << 7) | (mi["non_zero_dct"].astype(np.uint32) << 8))


def oracle_bpred_words(mi):
    b = mi.ravel()["bpred"].astype(np.uint32)
    lo = sum(b[:, i] << (4 * i) for i in range(8))
    hi = sum(b[:, 8 + i] << (4 * i) for i in range(8))
    return lo, hi


def _cases():
    out = []
    for n in ("1", "2", "gallery2_3_webp_a", "gallery2_1_webp_a", "regression_dark"):
        out.append((n, open(os.path.join(HERE, "golden", "decode", n + ".webp"), "rb").read(), None))
    img = PI.crop("3", 256, 104, 384, 256)
    for q, m in ((75, 4), (20, 4), (95, 6), (50, 0)):
        out.append(("enc q%d m%d" % (q, m), O.encode(img, q, m)[1], img))
    for (w, h) in ((99, 87), (17, 17), (1, 1), (300, 9), (33, 250)):
        im = synth.photo_like(w, h, 7)
        out.append(("synth %dx%d" % (w, h), O.encode(im, 60, 4)[1], im))
    from PIL import Image
    for q in (10, 50, 100):
        b = io.BytesIO()
        Image.fromarray(img).save(b, "WEBP", quality=q, method=4)
        out.append(("libwebp q%d" % q, b.getvalue(), img))
    return out


@pytest.mark.parametrize("name,data,src", _cases(), ids=[c[0] for c in _cases()])
def test_device_decoder_source_matches_oracle(name, data, src):
    for fancy in (True, False):
        rc, planes, mbinfo, rgb, st, sse = hc_decode(data, fancy, src)
        rco, o = O.decode(data, fancy, ("rgb", "planes", "mbinfo"))
        assert rc == 0 and rco == 0
        op = np.concatenate([o["planes"]["y"].ravel(), o["planes"]["u"].ravel(), o["planes"]["v"].ravel()])
        assert np.array_equal(oracle_flags(o["mbinfo"]), mbinfo[:, 0]), "modes / flags"
        lo, hi = oracle_bpred_words(o["mbinfo"])
        is_b = (mbinfo[:, 0] & 7) == 4
        assert np.array_equal(lo[is_b], mbinfo[is_b, 1]) and np.array_equal(hi[is_b], mbinfo[is_b, 2]), "sub-block modes"
        assert np.array_equal(op, planes), "filtered planes"
        assert np.array_equal(o["rgb"], rgb), "rgb"
        if src is not None:
            assert sse == int(((o["rgb"].astype(np.int64) - src.astype(np.int64)) ** 2).sum())
        assert st[1] == o["hdr"]["filter_type"] and st[2] == o["hdr"]["filter_level"]


def test_device_decoder_errors():
    data = O.encode(synth.photo_like(64, 48, 1), 75, 4, container=False)[1]
    for bad, want in ((data[:3] + b"\x9d\x01\x2b" + data[6:], 3), (bytes([data[0] | 1]) + data[1:], 2), (data[:40], None), (data[:len(data) // 2], None)):
        vp8 = bad
        planes = np.zeros(4 * 3 * 384, np.uint8)
        mbinfo = np.zeros((12, 4), np.uint32)
        rc = H.hc_decode(vp8, len(vp8), 64, 48, 1, planes.ctypes.data, mbinfo.ctypes.data, None, None, 0, None, None)
        rco = O.decode(bad)[0]
        assert rc == rco and rc != 0 and (want is None or rc == want)
    # the arenas are laid out for the dimensions the host read: a header that disagrees is refused, not decoded
    planes = np.zeros(4 * 3 * 384, np.uint8)
    mbinfo = np.zeros((12, 4), np.uint32)
    assert H.hc_decode(data, len(data), 64, 32, 1, planes.ctypes.data, mbinfo.ctypes.data, None, None, 0, None, None) == 7
