"""ctypes binding of the CPU oracle (oracle/_build/libzw_oracle.so) -- test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "libzw_oracle.so")


def build():
    src = os.path.join(ROOT, "oracle", "zw_oracle.cpp")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


class MbRecord(C.Structure):
    _fields_ = [("ymode", C.c_uint8), ("uvmode", C.c_uint8), ("segment", C.c_uint8), ("skip", C.c_uint8),
                ("bmodes", C.c_uint8 * 16), ("top_nz", C.c_uint16), ("left_nz", C.c_uint16),
                ("derr_left", C.c_int8 * 4), ("derr_top", C.c_int8 * 4), ("levels", (C.c_int16 * 16) * 25)]


MB_DTYPE = np.dtype([("ymode", "u1"), ("uvmode", "u1"), ("segment", "u1"), ("skip", "u1"),
                     ("bmodes", "u1", (16,)), ("top_nz", "<u2"), ("left_nz", "<u2"),
                     ("derr_left", "i1", (4,)), ("derr_top", "i1", (4,)), ("levels", "<i2", (25, 16))])
assert MB_DTYPE.itemsize == 832 == C.sizeof(MbRecord)

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        u8p = C.POINTER(C.c_uint8)
        L.zwo_encode_vp8.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(u8p), C.POINTER(C.c_size_t), C.c_void_p]
        L.zwo_encode_webp.argtypes = L.zwo_encode_vp8.argtypes
        L.zwo_free.argtypes = [C.c_void_p]
        L.zwo_dump_new.restype = C.c_void_p
        L.zwo_dump_free.argtypes = [C.c_void_p]
        L.zwo_dump_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(u8p), C.POINTER(C.c_size_t)]
        L.zwo_encode_batch_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.zwo_encode_batch_mt.restype = C.c_size_t
        L.zwo_encode_batch_mt_out.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_size_t, C.c_void_p]
        L.zwo_encode_batch_mt_out.restype = C.c_size_t
        L.zwo_cbrt.restype = C.c_double
        L.zwo_cbrt.argtypes = [C.c_double]
        L.zwo_pow.restype = C.c_double
        L.zwo_pow.argtypes = [C.c_double, C.c_double]
        L.zwo_bool_encode.restype = C.c_size_t
        L.zwo_bool_encode_tree.restype = C.c_size_t
        L.zwo_rd_score.restype = C.c_uint32
        L.zwo_residual_cost.restype = C.c_uint32
        for f in ("zwo_fixed_cost_i16", "zwo_fixed_cost_uv", "zwo_fixed_cost_i4", "zwo_entropy_cost", "zwo_level_fixed_cost"):
            getattr(L, f).restype = C.c_uint16
    return _lib


COLOR = {"L8": 0, "La8": 1, "Rgb8": 2, "Rgba8": 3}

STAGES = ["YUV_Y", "YUV_U", "YUV_V", "BASE_QIDX", "SEG_ENABLED", "ALPHA", "ALPHA_HIST", "SEG_CENTERS", "SEG_MAP256",
          "SEG_MID", "SEG_MAP", "SEG_QIDX", "SEG_TREE_PROBS", "SEG_UPDATE_MAP", "P1MB", "STATS", "PROBS",
          "PROBS_UPDATED", "SKIP_PROB", "LCOST", "P2MB", "HDR_TOKENS", "TOK_TOKENS", "PART0", "PART1", "TOKEN_PROBS_FINAL", "VP8"]
_STAGE_DTYPES = {"HDR_TOKENS": "<u2", "TOK_TOKENS": "<u2", "ALPHA_HIST": "<u4", "SEG_MID": "<i4", "STATS": "<u4", "LCOST": "<u2", "P1MB": MB_DTYPE, "P2MB": MB_DTYPE}


def encode(img, quality, method, color="Rgb8", container=True, want_dump=False):
    """img: uint8 array [h,w,c].  Returns (status, bytes, dump-dict-or-None)."""
    L = lib()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[0], img.shape[1]
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    d = L.zwo_dump_new() if want_dump else None
    fn = L.zwo_encode_webp if container else L.zwo_encode_vp8
    rc = fn(img.ctypes.data, img.nbytes, w, h, COLOR[color], int(quality), int(method), C.byref(out), C.byref(n), d)
    data = b""
    if rc == 0:
        data = C.string_at(out, n.value)
        L.zwo_free(out)
    dump = None
    if want_dump:
        dump = {}
        for name in STAGES:
            p = C.POINTER(C.c_uint8)()
            ln = C.c_size_t(0)
            if L.zwo_dump_get(d, name.encode(), C.byref(p), C.byref(ln)):
                raw = C.string_at(p, ln.value)
                dump[name] = np.frombuffer(raw, dtype=_STAGE_DTYPES.get(name, "u1")).copy()
        L.zwo_dump_free(d)
    return rc, data, dump


DEC_MB_DTYPE = np.dtype([("luma_mode", "u1"), ("chroma_mode", "u1"), ("segment", "u1"), ("skipped", "u1"), ("non_zero_dct", "u1"),
                         ("pad", "u1", (3,)), ("bpred", "u1", (16,))])
DEC_HDR = ["width", "height", "mbw", "mbh", "filter_type", "filter_level", "sharpness", "num_partitions", "segments_enabled",
           "update_map", "lf_adj", "has_skip_prob", "prob_skip_false", "version", "pixel_type", "reserved"]


def decode(data: bytes, fancy=True, want=("rgb",)):
    """Decode a VP8 key frame (bare or RIFF/WEBP) with the decoder oracle.  Returns (status, dict) with the keys asked
    for in `want` out of rgb [h,w,3], planes / planes_unfiltered (dict y,u,v of the padded frame), mbinfo, plus hdr."""
    L = lib()
    u8p = C.POINTER(C.c_uint8)
    L.zwo_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(u8p), C.POINTER(u8p), C.POINTER(u8p), C.POINTER(u8p), C.POINTER(C.c_uint32)]
    ptr = {k: u8p() for k in ("rgb", "planes", "planes_unfiltered", "mbinfo")}
    hdr = (C.c_uint32 * 16)()
    rc = L.zwo_decode(data, len(data), 1 if fancy else 0, *[C.byref(ptr[k]) if k in want else None for k in ("rgb", "planes", "planes_unfiltered", "mbinfo")], hdr)
    out = {"hdr": dict(zip(DEC_HDR, list(hdr)))}
    if rc != 0:
        return rc, out
    w, h, mbw, mbh = hdr[0], hdr[1], hdr[2], hdr[3]
    ysz, csz = mbw * mbh * 256, mbw * mbh * 64
    for k in want:
        if k == "rgb":
            out[k] = np.frombuffer(C.string_at(ptr[k], w * h * 3), np.uint8).reshape(h, w, 3).copy()
        elif k in ("planes", "planes_unfiltered"):
            raw = np.frombuffer(C.string_at(ptr[k], ysz + 2 * csz), np.uint8)
            out[k] = {"y": raw[:ysz].reshape(mbh * 16, mbw * 16).copy(), "u": raw[ysz:ysz + csz].reshape(mbh * 8, mbw * 8).copy(),
                      "v": raw[ysz + csz:].reshape(mbh * 8, mbw * 8).copy()}
        elif k == "mbinfo":
            out[k] = np.frombuffer(C.string_at(ptr[k], mbw * mbh * 24), DEC_MB_DTYPE).reshape(mbh, mbw).copy()
        L.zwo_free(ptr[k])
    return rc, out


def encode_batch_mt(imgs, quality, method, threads=None, color="Rgb8", container=True, L=None):
    """imgs: uint8 array [n,h,w,c] of same-sized images -> list of n files (bytes), encoded with `threads` host threads."""
    L = L or lib()
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    n, h, w = imgs.shape[0], imgs.shape[1], imgs.shape[2]
    threads = threads or (os.cpu_count() or 1)
    stride = 64 + 2 * w * h  # far above any real file (q100 noise stays below 1.6 B/px)
    arena = np.empty((n, stride), np.uint8)
    lens = np.zeros(n, np.uint32)
    L.zwo_encode_batch_mt_out(imgs.ctypes.data, n, w, h, COLOR[color], int(quality), int(method), int(threads), 1 if container else 0,
                              arena.ctypes.data, stride, lens.ctypes.data)
    assert (lens > 0).all() and (lens <= stride).all()
    return [arena[i, :lens[i]].tobytes() for i in range(n)]


def encode_raw(data: bytes, w, h, quality, method, color="Rgb8", container=True):
    L = lib()
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if data else b"\0")
    fn = L.zwo_encode_webp if container else L.zwo_encode_vp8
    rc = fn(buf, len(data), w, h, COLOR[color], int(quality), int(method), C.byref(out), C.byref(n), None)
    res = b""
    if rc == 0:
        res = C.string_at(out, n.value)
        L.zwo_free(out)
    return rc, res


# ---- algorithmic int-op model: oracle-counted primitive invocations x fixed costs ----
OP_NAMES = ["fdct", "idct", "wht", "iwht", "ttransform", "quant_coeff", "sse_px", "cost_coeff", "trellis_pos", "i4_predset",
            "add_residue", "trellis_block"]
# 32-bit integer operations per invocation, loads/stores excluded: FDCT incl. the residual 190, IDCT 176, WHT 112,
# IWHT 96, TTransform 112, quantise+dequantise 8 per coefficient, SSE 3 per pixel, residual cost 10 per coefficient
# visited, trellis 60 i64 ops (counted x2.5) per position, the ten 4x4 predictors 100 per set, add_residue 48.
OP_COST = [190, 176, 112, 96, 112, 8, 3, 10, 150, 100, 48, 0]
OP_SETS = ["pass1_luma", "pass1_chroma", "pass2_luma", "pass2_chroma"]


def count_ops(img, quality, method):
    """Encode `img` with the oracle on this thread; return {set: {primitive: count}} and {set: int ops}."""
    import ctypes as C
    L = lib()
    L.zwo_opcounts_reset()
    rc, _, _ = encode(img, quality, method)
    assert rc == 0
    buf = (C.c_uint64 * 64)()
    L.zwo_opcounts_get.restype = C.c_size_t
    n = L.zwo_opcounts_get(buf, 64)
    counts = {s: {OP_NAMES[i]: int(buf[k * n + i]) for i in range(n)} for k, s in enumerate(OP_SETS)}
    ops = {s: sum(counts[s][OP_NAMES[i]] * OP_COST[i] for i in range(n)) for s in OP_SETS}
    return counts, ops


# ---- VP8L lossless encoder + full container (oracle/zw_lossless_oracle.inc) ----
def _take(L, rc, out, n):
    if rc != 0:
        return rc, b""
    data = C.string_at(out, n.value)
    L.zwo_free(out)
    return 0, data


def encode_lossless(img, color="Rgb8", use_predictor=True, implicit_dimensions=False):
    """encode_frame_lossless: the raw VP8L stream of `img` (uint8 [h,w,c] or [h,w])."""
    L = lib()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[0], img.shape[1]
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    L.zwo_encode_lossless.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
    rc = L.zwo_encode_lossless(img.ctypes.data, img.nbytes, w, h, COLOR[color], int(use_predictor), int(implicit_dimensions),
                               C.byref(out), C.byref(n))
    return _take(L, rc, out, n)


def encode_alpha_lossless(img, color="Rgba8"):
    L = lib()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[0], img.shape[1]
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    L.zwo_encode_alpha_lossless.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int,
                                            C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
    rc = L.zwo_encode_alpha_lossless(img.ctypes.data, img.nbytes, w, h, COLOR[color], C.byref(out), C.byref(n))
    return _take(L, rc, out, n)


def webp_encode(img, color="Rgb8", use_predictor=True, use_lossy=False, quality=95, method=4, icc=b"", exif=b"", xmp=b"",
                raw=None, w=None, h=None):
    """WebPEncoder::encode: the complete .webp file (simple or VP8X container).  raw/w/h override img for bad-size tests."""
    L = lib()
    if raw is None:
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[0], img.shape[1]
        raw = img.tobytes()
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    L.zwo_webp_encode.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                  C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
    rc = L.zwo_webp_encode(raw, len(raw), w, h, COLOR[color], int(use_predictor), int(use_lossy), int(quality), int(method),
                           icc, len(icc), exif, len(exif), xmp, len(xmp), C.byref(out), C.byref(n))
    return _take(L, rc, out, n)


def build_huffman(freqs, limit):
    L = lib()
    f = np.ascontiguousarray(freqs, dtype=np.uint32)
    lengths = np.zeros(len(f), np.uint8)
    codes = np.zeros(len(f), np.uint16)
    L.zwo_build_huffman.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    ok = L.zwo_build_huffman(f.ctypes.data, len(f), int(limit), lengths.ctypes.data, codes.ctypes.data)
    return bool(ok), lengths, codes


def webp_encode_batch_mt(imgs, color="Rgb8", use_predictor=True, use_lossy=False, quality=95, method=4, threads=None, L=None, keep=True):
    """WebPEncoder::encode of n same-sized images ([n,h,w,c] uint8) on `threads` host threads -> (files or None, seconds)."""
    import time
    L = L or lib()
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    n, h, w = imgs.shape[0], imgs.shape[1], imgs.shape[2]
    threads = threads or (os.cpu_count() or 1)
    stride = 64 + 8 * w * h if keep else 0
    arena = np.empty((n, stride), np.uint8) if keep else None
    lens = np.zeros(n, np.uint32)
    L.zwo_webp_encode_batch_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_size_t, C.c_void_p]
    L.zwo_webp_encode_batch_mt.restype = C.c_size_t
    t0 = time.perf_counter()
    L.zwo_webp_encode_batch_mt(imgs.ctypes.data, n, w, h, COLOR[color], int(use_predictor), int(use_lossy), int(quality), int(method),
                               int(threads), arena.ctypes.data if keep else None, stride, lens.ctypes.data)
    dt = time.perf_counter() - t0
    assert (lens > 0).all()
    return ([arena[i, :lens[i]].tobytes() for i in range(n)] if keep else None), dt
