"""Pin the CPU oracle against every known-answer test the reference holds for the lossy encode
path.  Each test cites the reference test it restates."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

L = O.lib()


def _bits_literal(nbits, value):
    return [((1 << b) & value) > 0 for b in range(nbits - 1, -1, -1)]


def _bool_encode(ops):
    bits = (C.c_uint8 * len(ops))(*[int(b) for b, _ in ops])
    probs = (C.c_uint8 * len(ops))(*[p for _, p in ops])
    out = (C.c_uint8 * (len(ops) + 16))()
    n = L.zwo_bool_encode(bits, probs, len(ops), out)
    return bytes(out[:n])


def _short_ops():
    ops = [(False, 128), (True, 10), (False, 250)]
    ops += [(b, 128) for b in _bits_literal(1, 1)]
    ops += [(b, 128) for b in _bits_literal(3, 5)]
    ops += [(b, 128) for b in _bits_literal(8, 64)]
    ops += [(b, 128) for b in _bits_literal(8, 185)]
    return ops


def test_arithmetic_encoder_short():
    # src/encoder/arithmetic.rs:212-223
    assert _bool_encode(_short_ops()) == bytes([104, 101, 107, 128])


def test_arithmetic_encoder_hello():
    # src/encoder/arithmetic.rs:226-243
    ops = _short_ops()
    ops += [(b, 128) for b in _bits_literal(8, 31)]
    ops += [(b, 128) for b in _bits_literal(8, 134)]
    ops += [(False, 128)]  # write_optional_signed_value(2, None)
    ops += [(True, 128)] + [(b, 128) for b in _bits_literal(2, 1)] + [(True, 128)]  # Some(1): flag, |v|, sign>=0
    assert _bool_encode(ops)[:5] == b"hello"


def _bool_decode(data, probs):
    """RFC 6386 section 7.3 boolean decoder (independent of the oracle)."""
    data = bytes(data) + b"\0\0"
    value = (data[0] << 8) | data[1]
    pos, rng, bit_count, out = 2, 255, 0, []
    for p in probs:
        split = 1 + (((rng - 1) * p) >> 8)
        SPLIT = split << 8
        if value >= SPLIT:
            out.append(True)
            rng -= split
            value -= SPLIT
        else:
            out.append(False)
            rng = split
        while rng < 128:
            value = (value << 1) & 0xFFFFFF
            rng <<= 1
            bit_count += 1
            if bit_count == 8:
                bit_count = 0
                value |= data[pos] if pos < len(data) else 0
                pos += 1
    return out


def test_encoder_with_decoder():
    # src/encoder/arithmetic.rs:245-266
    ops = [(True, 40), (True, 110), (False, 70), (False, 10), (True, 5)]
    enc = _bool_encode(ops)
    assert _bool_decode(enc, [p for _, p in ops]) == [b for b, _ in ops]


def test_encoder_decoder_random_roundtrip():
    rng = np.random.default_rng(7)
    for n in (1, 7, 64, 5000):
        probs = rng.integers(1, 256, n).tolist()
        bits = [bool(rng.random() > p / 256.0) for p in probs]
        enc = _bool_encode(list(zip(bits, probs)))
        assert _bool_decode(enc, probs) == bits


def test_encoder_tree():
    # src/encoder/arithmetic.rs:268-274 : write_with_tree(KEYFRAME_YMODE_TREE, KEYFRAME_YMODE_PROBS, TM_PRED)
    probs = (C.c_uint8 * 4)(145, 156, 163, 128)
    vals = (C.c_int8 * 1)(3)
    out = (C.c_uint8 * 16)()
    n = L.zwo_bool_encode_tree(1, probs, vals, 1, 0, out)
    assert bytes(out[:n]) == bytes([233, 64, 0, 0])


def test_trellis_vs_libwebp():
    # src/encoder/cost.rs:2598-2675
    q = (C.c_uint16 * 16)(); iq = (C.c_uint32 * 16)(); bias = (C.c_uint32 * 16)()
    zt = (C.c_uint32 * 16)(); sh = (C.c_uint16 * 16)()
    L.zwo_matrix_new(25, 31, 0, q, iq, bias, zt, sh)
    assert list(q) == [25] + [31] * 15
    iq[0] = 5242
    for i in range(1, 16):
        iq[i] = 4228
    coeffs = (C.c_int32 * 16)(-282, 6, 3, -4, -3, -11, -4, -2, 5, 3, 4, -1, 2, -2, -3, -1)
    out = (C.c_int32 * 16)()
    probs = O.coeff_probs_default() if hasattr(O, "coeff_probs_default") else None
    import re, os
    txt = open(os.path.join(O.ROOT, "oracle", "vp8_tables.h")).read()
    m = re.search(r"kCoeffProbs\[1056\] = \{([^}]*)\}", txt)
    p = np.array([int(t) for t in m.group(1).replace("\n", "").split(",") if t.strip()], np.uint8)
    assert p.size == 1056
    L.zwo_trellis(coeffs, out, q, iq, bias, sh, 840, 0, p.ctypes.data_as(C.POINTER(C.c_uint8)), 3, 0)
    assert list(out) == [-11] + [0] * 15
    # libwebp debug log values quoted in the same test: skip_cost 89, init cost 576, thresh 240
    lc = np.zeros(4 * 8 * 3 * 68, np.uint16)
    L.zwo_level_costs(p.ctypes.data_as(C.POINTER(C.c_uint8)), lc.ctypes.data_as(C.POINTER(C.c_uint16)))
    p4 = p.reshape(4, 8, 3, 11)
    assert L.zwo_entropy_cost(int(p4[3, 0, 0, 0])) == 89          # eob cost = bit_cost(0, p0)
    assert L.zwo_entropy_cost(255 - int(p4[3, 0, 0, 0])) == 576   # init cost = bit_cost(1, p0)
    assert (31 * 31) // 4 == 240


def test_table_spot_values():
    # src/encoder/cost.rs:2036-2068, :1987-1997
    assert [L.zwo_fixed_cost_i16(i) for i in range(4)] == [663, 919, 872, 919]
    assert [L.zwo_fixed_cost_uv(i) for i in range(4)] == [302, 984, 439, 642]
    assert L.zwo_fixed_cost_i4(0, 0, 0) < 100 < L.zwo_fixed_cost_i4(0, 0, 1)
    assert abs(L.zwo_entropy_cost(128) - 256) < 10
    assert L.zwo_entropy_cost(255) < 10
    assert L.zwo_entropy_cost(1) > 1500
    assert L.zwo_level_fixed_cost(0) == 0


def test_lambda_formulas_and_quant_indices():
    # src/encoder/cost.rs:2089-2107 (formulas), SURVEY.md §8 (q50/q75/q90 values), vp8.rs:37-55
    assert L.zwo_quality_to_quant_index(75) == 26
    assert L.zwo_quality_to_quant_index(50) == 39
    assert L.zwo_quality_to_quant_index(90) == 9
    assert L.zwo_quality_to_quant_index(0) == 127
    assert L.zwo_quality_to_quant_index(100) == 0
    lam = (C.c_uint32 * 8)(); qs = (C.c_int16 * 6)()
    L.zwo_segment_lambdas(26, lam, qs)
    assert list(qs) == [24, 30, 48, 46, 24, 30]
    assert list(lam) == [21, 6348, 42, 7, 787, 529, 1800, 46]
    L.zwo_segment_lambdas(39, lam, qs)
    assert list(qs)[:4] == [36, 43, 72, 66]
    assert list(lam)[:6] == [43, 13068, 86, 14, 1617, 1089] and lam[7] == 67
    L.zwo_segment_lambdas(9, lam, qs)
    assert list(qs)[:4] == [12, 13, 24, 20]
    assert list(lam)[:6] == [3, 1200, 7, 1, 147, 100] and lam[7] == 20
    # generic formula check for q=64-like relationship: lambda_i16 = 3 q^2, i4 = 3q^2>>7, uv = 3q^2>>6
    for idx in range(0, 128, 7):
        L.zwo_segment_lambdas(idx, lam, qs)
        q_i4 = (qs[0] + 15 * qs[1] + 8) >> 4
        q_i16 = (qs[2] + 15 * qs[3] + 8) >> 4
        q_uv = (qs[4] + 15 * qs[5] + 8) >> 4
        assert lam[0] == max(1, (3 * q_i4 * q_i4) >> 7)
        assert lam[1] == max(1, 3 * q_i16 * q_i16)
        assert lam[2] == max(1, (3 * q_uv * q_uv) >> 6)
        assert lam[3] == max(1, (q_i4 * q_i4) >> 7)


def test_rd_score():
    # src/encoder/cost.rs:2122-2135
    full = C.c_uint64()
    L.zwo_rd_score(0, 0, 106, C.byref(full)); assert full.value == 0
    L.zwo_rd_score(100, 0, 106, C.byref(full)); assert full.value == 100 * 256
    L.zwo_rd_score(0, 663, 106, C.byref(full)); assert full.value == 663 * 106
    L.zwo_rd_score(1000, 663, 106, C.byref(full)); assert full.value == 1000 * 256 + 663 * 106


def test_dct_inverse():
    # src/common/transform.rs:214-228 + the FDCT anchor computed in SURVEY.md §8c
    block = [38, 6, 210, 107, 42, 125, 185, 151, 241, 224, 125, 233, 227, 8, 57, 96]
    b = (C.c_int32 * 16)(*block)
    L.zwo_dct4x4(b)
    assert list(b) == [1037, -83, 97, 130, -104, -290, -280, 101, -289, 27, 89, 235, 202, 63, 69, -69]
    L.zwo_idct4x4(b)
    assert list(b) == block


def test_wht_roundtrip_small():
    rng = np.random.default_rng(3)
    for _ in range(50):
        v = (rng.integers(-255, 256, 16) * 8).astype(np.int32)  # multiples of 8 survive (x+3)>>3 exactly
        b = (C.c_int32 * 16)(*v.tolist())
        L.zwo_wht4x4(b)
        L.zwo_iwht4x4(b)
        # WHT halves, IWHT >>3 : forward*inverse = 16/2/8 = 1
        assert np.abs(np.array(list(b)) - v).max() <= 1


def _pad(img, stride):
    buf = np.zeros(len(img) + 2 * stride + 16, np.uint8)
    buf[stride:stride + len(img)] = img
    return buf


def _predict(img, mode, x0, y0, stride):
    buf = _pad(np.array(img, np.uint8), stride)
    ptr = C.c_void_p(buf.ctypes.data + stride)
    L.zwo_predict4x4(ptr, mode, x0, y0, stride)
    return buf[stride:stride + len(img)].tolist()


def test_add_residue():
    # src/common/prediction.rs:959-971
    p = (C.c_uint8 * 16)(*range(1, 17))
    r = (C.c_int32 * 16)(-1, -2, -3, -4, 250, 249, 248, 250, -10, -18, -192, -17, -3, 15, 18, 9)
    L.zwo_add_residue(p, r, 0, 0, 4)
    assert list(p) == [0, 0, 0, 0, 255, 255, 255, 255, 0, 0, 0, 0, 10, 29, 33, 25]


def test_predict_bhepred():
    # src/common/prediction.rs:974-993
    im = [5, 0, 0, 0, 0, 4, 0, 0, 0, 0, 3, 0, 0, 0, 0, 2, 0, 0, 0, 0, 1, 0, 0, 0, 0]
    exp = [5, 0, 0, 0, 0, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1]
    assert _predict(im, 3, 1, 1, 5) == exp


def test_predict_brdpred():
    # src/common/prediction.rs:995-1014
    im = [5, 6, 7, 8, 9, 4, 0, 0, 0, 0, 3, 0, 0, 0, 0, 2, 0, 0, 0, 0, 1, 0, 0, 0, 0]
    exp = [5, 6, 7, 8, 9, 4, 5, 6, 7, 8, 3, 4, 5, 6, 7, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5]
    assert _predict(im, 5, 1, 1, 5) == exp


def test_predict_bldpred():
    # src/common/prediction.rs:1016-1054
    im = [1, 2, 3, 4, 5, 6, 7, 8] + [0] * 64
    out = _predict(im, 4, 0, 1, 8)
    a = [2, 3, 4, 5, 6, 7, 8]
    for y in range(4):
        assert out[8 * (y + 1):8 * (y + 1) + 4] == a[y:y + 4]


def test_predict_bvepred():
    # src/common/prediction.rs:1056-1090
    im = [1, 2, 3, 4, 5, 6, 7, 8, 9] + [0] * 72
    out = _predict(im, 2, 1, 1, 9)
    for y in range(4):
        assert out[9 * (y + 1) + 1:9 * (y + 1) + 5] == [2, 3, 4, 5]


def test_avg_helpers_via_predictors():
    # exhaustive avg2 / sampled avg3 (prediction.rs:863-914) through HU (avg2(l0,l1) at [0]) and
    # VE (avg3(p,a0,a1) at [0]).
    stride = 8
    for i in range(0, 256, 3):
        for j in range(0, 256, 5):
            im = [0] * (stride * 5)
            im[stride * 1 + 0] = i  # l0 (x0=1,y0=1 -> left col at x=0)
            im[stride * 2 + 0] = j  # l1
            out = _predict(im, 9, 1, 1, stride)
            assert out[stride + 1] == (i + j + 1) // 2
    rng = np.random.default_rng(5)
    for _ in range(3000):
        p, a0, a1 = [int(v) for v in rng.integers(0, 256, 3)]
        im = [0] * (stride * 5)
        im[0], im[1], im[2] = p, a0, a1
        out = _predict(im, 2, 1, 1, stride)
        assert out[stride + 1] == (p + 2 * a0 + a1 + 2) // 4


def test_i4_predictions_match_inplace_predictors():
    rng = np.random.default_rng(11)
    stride = 32
    for _ in range(20):
        ws = rng.integers(0, 256, stride * 17, dtype=np.uint8)
        allp = (C.c_uint8 * 160)()
        buf = _pad(ws, stride)
        L.zwo_predict4x4_all(C.c_void_p(buf.ctypes.data + stride), 5, 5, stride, allp)
        for m in range(10):
            out = np.array(_predict(ws.tolist(), m, 5, 5, stride), np.uint8).reshape(17, stride)
            assert out[5:9, 5:9].reshape(-1).tolist() == list(allp[m * 16:(m + 1) * 16])


def test_fast_math():
    # src/encoder/fast_math.rs:128-206 (cbrt < 1e-10 rel, pow < 1 %)
    L.zwo_cbrt.restype = C.c_double
    for x in (0.001, 0.1, 0.3333, 0.5, 0.9, 1.0):
        assert abs(L.zwo_cbrt(x) - x ** (1 / 3)) < 1e-10
    for x, n in ((0.5, 1.2), (0.8, 0.9), (0.2, 1.35), (0.93, 0.65)):
        assert abs(L.zwo_pow(x, n) - x ** n) / x ** n < 0.01


def test_record_coeffs_quirks():
    # cost.rs:1297-1397 : all-zero block records a single EOB=0 at node 0; Q10 skip_eob stays set
    st = (C.c_uint32 * 1056)()
    z = (C.c_int32 * 16)()
    L.zwo_record_coeffs(z, 3, 0, 0, st)
    s = np.array(list(st)).reshape(4, 8, 3, 11)
    assert s[3, 0, 0, 0] == 0x00010000 and s.sum() == 0x00010000
    st = (C.c_uint32 * 1056)()
    z = (C.c_int32 * 16)(0, 3, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    L.zwo_record_coeffs(z, 3, 0, 1, st)
    s = np.array(list(st)).reshape(4, 8, 3, 11)
    # pos0 (band0, ctx1): node0=1, node1=0 ; pos1 (band1, ctx0): skip_eob -> no node0; node1=1,node2=1,node3=0,node4=1,node5=0
    assert s[3, 0, 1, 0] == 0x00010001 and s[3, 0, 1, 1] == 0x00010000
    assert s[3, 1, 0, 0] == 0 and s[3, 1, 0, 1] == 0x00010001 and s[3, 1, 0, 5] == 0x00010000
    # pos2 (band2, ctx2): node0 NOT recorded because skip_eob is never cleared (Q10)
    assert s[3, 2, 2, 0] == 0 and s[3, 2, 2, 2] == 0x00010000
    # trailing EOB at pos3 (band 3) with ctx 1
    assert s[3, 3, 1, 0] == 0x00010000


def test_yuv_formula_and_padding():
    # decoder/yuv.rs:656-899
    rng = np.random.default_rng(1)
    w, h = 19, 5
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    mbw, mbh = 2, 1
    y = np.zeros(16 * mbw * 16 * mbh, np.uint8); u = np.zeros(8 * mbw * 8 * mbh, np.uint8); v = np.zeros_like(u)
    L.zwo_convert_yuv(img.ctypes.data_as(C.c_void_p), w, h, 3, y.ctypes.data_as(C.c_void_p),
                      u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
    y = y.reshape(16, 32); u = u.reshape(8, 16); v = v.reshape(8, 16)
    r, g, b = [img[:, :, i].astype(np.int64) for i in range(3)]
    yy = (16839 * r + 33059 * g + 6420 * b + 32768 + (16 << 16)) >> 16
    assert (y[:h, :w] == yy).all()
    assert (y[:h, w:] == y[:h, w - 1:w]).all() and (y[h:, :] == y[h - 1, :]).all()
    # chroma: 2x2 average with edge duplication
    ur = -9719 * r - 19081 * g + 28800 * b + (128 << 16)
    idx_r = np.minimum(np.arange(6), h - 1)
    idx_c = np.minimum(np.arange(20), w - 1)
    urp = ur[idx_r][:, idx_c]
    uavg = (urp[0::2, 0::2] + urp[0::2, 1::2] + urp[1::2, 0::2] + urp[1::2, 1::2] + (32768 << 2)) >> 18
    assert (u[:3, :10] == uavg).all()
    assert (u[:3, 10:] == u[:3, 9:10]).all() and (u[3:, :] == u[2, :]).all()


def test_symbol_log_reproduces_the_partitions():
    """The dumped (bit, probability) streams, pushed through the oracle's stand-alone boolean coder, give the
    dumped partitions back: pins the symbol dump the GPU tokeniser is compared against."""
    from image_webp_b200 import synth
    for img, q, m in ((synth.photo_like(96, 64, 3), 75, 4), (synth.noise(48, 48, 1), 90, 6), (synth.solid(32, 32), 50, 0)):
        rc, ref, d = O.encode(img, q, m, want_dump=True)
        assert rc == 0
        for sym, part in (("HDR_TOKENS", "PART0"), ("TOK_TOKENS", "PART1")):
            s = d[sym]
            bits = np.ascontiguousarray((s >> 8).astype(np.uint8))
            probs = np.ascontiguousarray((s & 255).astype(np.uint8))
            out = np.zeros(s.size + 16, np.uint8)
            L.zwo_bool_encode.restype = C.c_size_t
            n = L.zwo_bool_encode(bits.ctypes.data_as(C.c_void_p), probs.ctypes.data_as(C.c_void_p), C.c_size_t(s.size),
                                  out.ctypes.data_as(C.c_void_p))
            assert bytes(out[:n]) == bytes(d[part]), (sym, q, m)


def test_op_counters_are_consistent():
    """Primitive counters (measurement only): every coded macroblock runs 16 luma + 8 chroma forward DCTs in the
    final transform of each pass, trellis blocks only exist in pass 2 at method >= 4, and the counters are per thread."""
    from image_webp_b200 import synth
    img = synth.photo_like(64, 48, 5)
    nmb = 4 * 3
    counts, ops = O.count_ops(img, 75, 4)
    assert counts["pass1_luma"]["trellis_block"] == 0 and counts["pass2_luma"]["trellis_block"] > 0
    assert counts["pass1_chroma"]["fdct"] >= 8 * nmb and counts["pass2_chroma"]["fdct"] >= 8 * nmb
    assert counts["pass1_luma"]["fdct"] >= 16 * nmb
    assert all(v > 0 for v in ops.values())
    counts0, _ = O.count_ops(img, 75, 0)
    assert counts0["pass2_luma"]["trellis_block"] == 0 and counts0["pass1_luma"]["i4_predset"] == 0
