"""CPU check of the per-lane CUDA primitives: the same headers the kernels use
(image_webp_b200/csrc/zw_prims.cuh, zw_cost.cuh) are compiled with g++ and compared with the
oracle on random inputs.  Catches arithmetic slips before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = O.ROOT
SRC = os.path.join(ROOT, "tests", "hostcheck", "zw_hostcheck.cpp")
SO = os.path.join(ROOT, "tests", "hostcheck", "_build", "libzw_hostcheck.so")


def _build():
    deps = [SRC] + [os.path.join(ROOT, "image_webp_b200", "csrc", f) for f in ("zw_prims.cuh", "zw_cost.cuh", "zw_tables.inc", "zw_boolcoder.cuh", "zw_quad.cuh", "zw_types.cuh", "zw_dec.cuh", "zw_lossless.cuh")]
    if (not os.path.exists(SO)) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", SO])
    return C.CDLL(SO)


H = _build()
L = O.lib()
H.hc_residual_cost.restype = C.c_uint32
i32p = C.POINTER(C.c_int32)


def _probs(rng, random=False):
    import re
    txt = open(os.path.join(ROOT, "oracle", "vp8_tables.h")).read()
    m = re.search(r"kCoeffProbs\[1056\] = \{([^}]*)\}", txt)
    p = np.array([int(t) for t in m.group(1).replace("\n", "").split(",") if t.strip()], np.uint8)
    if random:
        p = rng.integers(1, 256, 1056).astype(np.uint8)
    return p


def test_transforms_match_oracle():
    rng = np.random.default_rng(0)
    for name in ("fdct", "idct", "wht", "iwht"):
        for _ in range(300):
            if name == "fdct":
                v = rng.integers(-255, 256, 16)
            elif name == "idct":
                v = rng.integers(-2100, 2101, 16)
                v[rng.random(16) < 0.5] = 0
            elif name == "wht":
                v = rng.integers(-2100, 2101, 16)
            else:
                v = rng.integers(-17000, 17001, 16)
            a = (C.c_int32 * 16)(*v.tolist()); b = (C.c_int32 * 16)(*v.tolist())
            getattr(H, "hc_" + name)(a)
            getattr(L, "zwo_" + ("dct" if name == "fdct" else name) + "4x4")(b)
            assert list(a) == list(b), name


def test_predictors_match_oracle():
    rng = np.random.default_rng(1)
    stride = 32
    for _ in range(200):
        ws = rng.integers(0, 256, stride * 17, dtype=np.uint8)
        if rng.random() < 0.2:
            ws[:] = rng.integers(0, 256)
        buf = np.zeros(stride * 19 + 16, np.uint8); buf[stride:stride + ws.size] = ws
        allp = (C.c_uint8 * 160)()
        x0, y0 = 1 + 4 * int(rng.integers(0, 4)), 1 + 4 * int(rng.integers(0, 4))
        L.zwo_predict4x4_all(C.c_void_p(buf.ctypes.data + stride), x0, y0, stride, allp)
        w2 = ws.reshape(17, stride)
        e = np.array([w2[y0 + 3, x0 - 1], w2[y0 + 2, x0 - 1], w2[y0 + 1, x0 - 1], w2[y0, x0 - 1]] +
                     [w2[y0 - 1, x0 - 1 + k] for k in range(9)], np.uint8)
        for m in range(10):
            out = (C.c_uint8 * 16)()
            H.hc_predict4(e.ctypes.data_as(C.c_void_p), m, out)
            assert list(out) == list(allp[m * 16:(m + 1) * 16]), m
            out2 = (C.c_uint8 * 16)()  # the two-level lookup form the cooperative I4 search uses
            H.hc_predict4_lut(e.ctypes.data_as(C.c_void_p), m, out2)
            assert list(out2) == list(allp[m * 16:(m + 1) * 16]), ("lut", m)


def test_ttransform_matches_oracle_tdisto():
    rng = np.random.default_rng(2)
    for _ in range(50):
        a = rng.integers(0, 256, (16, 16), dtype=np.uint8); b = rng.integers(0, 256, (16, 16), dtype=np.uint8)
        ref = L.zwo_tdisto_16x16(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 16)
        tot = 0
        for by in range(4):
            for bx in range(4):
                pa = a[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4].astype(np.int32).reshape(-1)
                pb = b[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4].astype(np.int32).reshape(-1)
                ta = H.hc_ttransform(pa.ctypes.data_as(i32p)); tb = H.hc_ttransform(pb.ctypes.data_as(i32p))
                tot += abs(tb - ta) >> 5
        assert tot == ref


def _matrix(qdc, qac, typ):
    q = (C.c_uint16 * 16)(); iq = (C.c_uint32 * 16)(); bias = (C.c_uint32 * 16)(); zt = (C.c_uint32 * 16)(); sh = (C.c_uint16 * 16)()
    L.zwo_matrix_new(qdc, qac, typ, q, iq, bias, zt, sh)
    return q, iq, bias, sh


def _levels(rng):
    lv = np.zeros(16, np.int32)
    n = int(rng.integers(0, 17))
    mag = rng.choice([1, 2, 3, 5, 12, 40, 80, 300, 2047])
    lv[:n] = rng.integers(-mag, mag + 1, n)
    lv[rng.random(16) < 0.3] = 0
    return lv


def test_residual_cost_and_level_costs_match_oracle():
    rng = np.random.default_rng(3)
    for trial in range(6):
        p = _probs(rng, random=trial > 0)
        lc = np.zeros(6528, np.uint16); lc2 = np.zeros(6528, np.uint16)
        L.zwo_level_costs(p.ctypes.data_as(C.c_void_p), lc.ctypes.data_as(C.c_void_p))
        H.hc_level_costs(p.ctypes.data_as(C.c_void_p), lc2.ctypes.data_as(C.c_void_p))
        assert (lc == lc2).all()
        for _ in range(400):
            lv = _levels(rng)
            ctype = int(rng.integers(0, 4)); first = 1 if ctype == 0 else 0; ctx0 = int(rng.integers(0, 3))
            if first == 1:
                lv[0] = 0
            for zero_tables in (0, 1):
                ref = L.zwo_residual_cost(lv.ctypes.data_as(i32p), ctype, first, ctx0, p.ctypes.data_as(C.c_void_p), zero_tables)
                got = H.hc_residual_cost(lv.ctypes.data_as(i32p), ctype, first, ctx0, p.ctypes.data_as(C.c_void_p),
                                         None if zero_tables else lc.ctypes.data_as(C.c_void_p))
                assert ref == got, (lv, ctype, first, ctx0, zero_tables)


def test_trellis_matches_oracle():
    rng = np.random.default_rng(4)
    for trial in range(2000):
        qi = int(rng.integers(0, 128))
        lam = (C.c_uint32 * 8)(); qs = (C.c_int16 * 6)()
        L.zwo_segment_lambdas(qi, lam, qs)
        q, iq, bias, sh = _matrix(qs[0], qs[1], 0)
        p = _probs(rng, random=(trial % 3 != 0))
        lc = np.zeros(6528, np.uint16)
        L.zwo_level_costs(p.ctypes.data_as(C.c_void_p), lc.ctypes.data_as(C.c_void_p))
        i4 = bool(rng.integers(0, 2))
        first, ctype, lamb = (0, 3, lam[4]) if i4 else (1, 0, lam[5])
        scale = rng.choice([3, 20, 100, 600, 2000])
        co = rng.integers(-scale, scale + 1, 16).astype(np.int32)
        co[rng.random(16) < 0.3] = 0
        ctx0 = int(rng.integers(0, 3))
        a = co.copy(); b = co.copy()
        oa = np.zeros(16, np.int32); ob = np.zeros(16, np.int32)
        q2 = (C.c_uint16 * 2)(q[0], q[1]); iq2 = (C.c_uint32 * 2)(iq[0], iq[1]); b2 = (C.c_uint32 * 2)(bias[0], bias[1])
        ra = L.zwo_trellis(a.ctypes.data_as(i32p), oa.ctypes.data_as(i32p), q, iq, bias, sh, lamb, first, p.ctypes.data_as(C.c_void_p), ctype, ctx0)
        rb = H.hc_trellis(b.ctypes.data_as(i32p), ob.ctypes.data_as(i32p), q2, iq2, b2, sh, lamb, first, p.ctypes.data_as(C.c_void_p),
                          lc.ctypes.data_as(C.c_void_p), ctype, ctx0)
        assert ra == rb and (oa == ob).all() and (a == b).all(), (trial, co, oa, ob)
        # the rolled form the quad kernels call (zw_quad.cuh q_trellis)
        c3 = co.copy(); o3 = np.zeros(16, np.int32)
        rc3 = H.hc_trellis_rolled(c3.ctypes.data_as(i32p), o3.ctypes.data_as(i32p), q2, iq2, b2, sh, lamb, first, p.ctypes.data_as(C.c_void_p),
                                  lc.ctypes.data_as(C.c_void_p), ctype, ctx0)
        assert rc3 == ra and (o3 == oa).all() and (c3 == a).all(), ("rolled", trial, co, oa, o3)


def test_token_events_match_oracle_record_coeffs():
    rng = np.random.default_rng(5)
    for _ in range(1500):
        lv = _levels(rng)
        t = int(rng.integers(0, 4)); first = 1 if t == 0 else 0; ctx = int(rng.integers(0, 3))
        if first:
            lv[0] = 0
        ref = (C.c_uint32 * 1056)()
        L.zwo_record_coeffs(lv.ctypes.data_as(i32p), t, first, ctx, ref)
        got = np.zeros(1056, np.uint32)
        z = lv.astype(np.int16)
        H.hc_token_events(z.ctypes.data_as(C.c_void_p), t, first, ctx, got.ctypes.data_as(C.c_void_p))
        assert (np.array(list(ref), np.uint32) == got).all()


def _ref_bool_encode(tok):
    bits = (tok >> 8).astype(np.uint8)
    probs = (tok & 255).astype(np.uint8)
    out = np.zeros(tok.size + 16, np.uint8)
    n = L.zwo_bool_encode(bits.ctypes.data_as(C.c_void_p), probs.ctypes.data_as(C.c_void_p), C.c_size_t(tok.size), out.ctypes.data_as(C.c_void_p))
    return out[:n].tobytes()


def _seg_bool_encode(tok, seg, warm):
    H.hc_boolcode_segmented.restype = C.c_size_t
    H.hc_boolcode_segmented.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    tok = np.ascontiguousarray(tok, np.uint16)
    out = np.full(tok.size + 64, 0xEE, np.uint8)
    stats = np.zeros(4, np.uint32)
    n = H.hc_boolcode_segmented(tok.ctypes.data, tok.size, seg, warm, out.ctypes.data, out.size, stats.ctypes.data)
    assert n != 2 ** 64 - 1
    return out[:n].tobytes(), stats


def test_segment_parallel_boolcoder_equals_the_serial_coder():
    """The five-step segment-parallel scheme the GPU runs (same lane-local code, zw_boolcoder.cuh) reproduces the
    reference's serial ArithmeticEncoder byte for byte: real token streams, adversarial streams (carry chains across
    segment starts, literal runs that keep all range states apart, empty / one-symbol streams), many segment sizes."""
    rng = np.random.default_rng(7)
    import photo_inputs as PI
    streams = []
    _, _, d = O.encode(PI.crop("3", 300, 200, 256, 192), 75, 4, want_dump=True)
    streams += [d["TOK_TOKENS"], d["HDR_TOKENS"]]
    _, _, d = O.encode(PI.crop("5", 100, 100, 128, 128), 100, 4, want_dump=True)
    streams += [d["TOK_TOKENS"]]
    n = 40000
    streams.append(((rng.integers(0, 2, n) << 8) | rng.integers(1, 256, n)).astype(np.uint16))        # uniform random
    streams.append(((np.ones(n, np.int64) << 8) | rng.integers(1, 4, n)).astype(np.uint16))           # improbable ones: long 0xFF / carry chains
    streams.append(((rng.integers(0, 2, n) << 8) | 128).astype(np.uint16))                            # literals: states stay apart
    streams.append(((rng.random(n) < 0.02).astype(np.int64) << 8 | 250).astype(np.uint16))            # nearly no shifts per symbol
    streams.append(((rng.random(n) < 0.98).astype(np.int64) << 8 | 5).astype(np.uint16))
    streams += [np.zeros(0, np.uint16), np.array([0x0180], np.uint16), np.array([0x0080] * 23, np.uint16)]
    worst = 0
    for tok in streams:
        ref = _ref_bool_encode(np.ascontiguousarray(tok, np.uint16))
        for seg, warm in ((8192, 1024), (512, 64), (64, 16), (8, 8), (16, 0)):
            got, st = _seg_bool_encode(tok, seg, warm)
            assert got == ref, "stream of %d symbols, seg %d warm %d: differs (%d vs %d bytes)" % (tok.size, seg, warm, len(got), len(ref))
            worst = max(worst, int(st[1]))
    assert worst >= 16  # the literal stream really keeps many candidate states alive: the general path is exercised


def _quad_luma(img, q, m, dump, pas):
    """Run zw_quad.cuh's four-lanes-per-macroblock luma path over the image on the CPU (same source the kernel compiles)."""
    h, w = img.shape[:2]
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    y = np.ascontiguousarray(dump["YUV_Y"])
    assert y.size == mbw * mbh * 256
    segmap = np.ascontiguousarray(dump["SEG_MAP"]) if "SEG_MAP" in dump and dump["SEG_ENABLED"][0] else None
    segq = np.ascontiguousarray(dump["SEG_QIDX"]) if "SEG_QIDX" in dump else np.zeros(4, np.uint8)
    probs = np.ascontiguousarray(dump["PROBS"])
    lcost = np.ascontiguousarray(dump["LCOST"])
    p2 = dump["P2MB"]
    uvnz = np.zeros(mbw * mbh, np.uint8)
    nzb = (p2["levels"][:, 17:25, :] != 0).any(axis=2)
    for b in range(8):
        uvnz |= (nzb[:, b].astype(np.uint8) << b)
    out = np.zeros(mbw * mbh, O.MB_DTYPE)
    H.hc_quad_luma_image(y.ctypes.data_as(C.c_void_p), mbw, mbh, pas, m, int(dump["BASE_QIDX"][0]),
                         segmap.ctypes.data_as(C.c_void_p) if segmap is not None else None, segq.ctypes.data_as(C.c_void_p),
                         probs.ctypes.data_as(C.c_void_p), lcost.ctypes.data_as(C.c_void_p), uvnz.ctypes.data_as(C.c_void_p),
                         out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.parametrize("name,q,m", [("synth", 75, 4), ("photo", 75, 4), ("photo", 50, 6), ("synth", 90, 2), ("photo", 75, 0), ("small", 75, 5),
                                      ("photo", 20, 3)])
def test_quad_luma_path_matches_oracle_records(name, q, m):
    """zw_quad.cuh (lane-private 4x4 work, four lanes per macroblock row) against the oracle's P1MB / P2MB: modes, sub-block
    modes, every coded luma level, and in pass 2 the skip flags and complexity contexts."""
    import photo_inputs as PI
    from image_webp_b200 import synth
    img = {"synth": lambda: synth.photo_like(320, 272, 21), "photo": lambda: PI.crop("3", 400, 200, 320, 272),
           "small": lambda: synth.photo_like(99, 87, 2)}[name]()
    rc, _, dump = O.encode(img, q, m, want_dump=True)
    assert rc == 0
    for pas, key in ((1, "P1MB"), (2, "P2MB")):
        got = _quad_luma(img, q, m, dump, pas)
        ref = dump[key]
        assert (got["ymode"] == ref["ymode"]).all(), "pass %d ymode: %s" % (pas, np.nonzero(got["ymode"] != ref["ymode"])[0][:8])
        assert (got["bmodes"] == ref["bmodes"]).all(), "pass %d bmodes differ at MB %s" % (pas, np.nonzero((got["bmodes"] != ref["bmodes"]).any(axis=1))[0][:8])
        if pas == 2:
            for f in ("skip", "top_nz", "left_nz"):
                assert (got[f] == ref[f]).all(), "pass 2 %s differs at MB %s" % (f, np.nonzero(got[f] != ref[f])[0][:8])
            lv_ref = ref["levels"][:, :17, :]
        else:
            # pass-1 records of skipped macroblocks are zeroed later (k_finish1); compare the unskipped ones
            keep = ref["skip"] == 0
            got, ref = got[keep], ref[keep]
            lv_ref = ref["levels"][:, :17, :]
        bad = np.nonzero((got["levels"][:, :17, :] != lv_ref).any(axis=(1, 2)))[0]
        assert bad.size == 0, "pass %d: luma levels differ at %d MBs, first %s" % (pas, bad.size, bad[:8])
