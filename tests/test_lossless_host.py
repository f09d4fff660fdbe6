"""CPU check of the lossless device source: image_webp_b200/csrc/zw_lossless.cuh -- the per-pixel transforms, token
rules, Huffman construction (heap order, length limiting, canonical codes), tree serialisation and header the kernels
call -- compiled with g++ (tests/hostcheck) and compared with the lossless oracle byte for byte.  The kernels' scans and
bit packing are checked on the GPU (tests/test_gpu_lossless.py)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from image_webp_b200 import synth
from test_device_prims_host import H

H.hc_lossless.restype = C.c_size_t
H.hc_lossless.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p,
                          C.c_void_p, C.c_void_p, C.c_void_p]
H.hc_ll_huffman.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
FLAG_PRED, FLAG_IMPLICIT, FLAG_ALPHA_PLANE = 1, 2, 4
BPP = {"L8": 1, "La8": 2, "Rgb8": 3, "Rgba8": 4}


def hc_lossless(img, color, flags, coded_color=None):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    cap = 64 + h * w * 12
    out = np.zeros(cap, np.uint8)
    n = H.hc_lossless(img.ctypes.data, w, h, BPP[color], O.COLOR[coded_color or color], flags, out.ctypes.data, cap, None, None, None, None)
    assert n <= cap
    return out[:n].tobytes()


def _images():
    rng = np.random.default_rng(3)
    yield "noise", rng.integers(0, 256, (64, 64, 4), dtype=np.uint8)
    yield "photo", np.dstack([synth.photo_like(99, 87, 1), (np.add.outer(np.arange(87), np.arange(99)) % 256).astype(np.uint8)])
    flat = np.full((70, 300, 4), 9, np.uint8)
    flat[30:, 100:] = 200
    flat[5, 5] = 1
    yield "flat-runs", flat
    yield "one", rng.integers(0, 256, (1, 1, 4), dtype=np.uint8)
    yield "row", rng.integers(0, 3, (1, 5000, 4), dtype=np.uint8)
    yield "col", rng.integers(0, 2, (5000, 1, 4), dtype=np.uint8)


@pytest.mark.parametrize("name,rgba", list(_images()), ids=[n for n, _ in _images()])
def test_device_source_matches_oracle(name, rgba):
    views = {"Rgba8": rgba, "Rgb8": rgba[:, :, :3], "La8": rgba[:, :, 1:3], "L8": rgba[:, :, 2]}
    for color, img in views.items():
        for pred in (True, False):
            rc, ref = O.encode_lossless(img, color, use_predictor=pred)
            assert rc == 0
            got = hc_lossless(img, color, FLAG_PRED if pred else 0)
            assert got == ref, (name, color, pred)
    for color in ("Rgba8", "La8"):  # ALPH payload: alpha plane as L8, predictor on, implicit dimensions
        rc, ref = O.encode_alpha_lossless(views[color], color)
        assert rc == 0
        got = hc_lossless(views[color], color, FLAG_PRED | FLAG_IMPLICIT | FLAG_ALPHA_PLANE, coded_color="L8")
        assert b"\x01" + got == ref, (name, color)


def test_huffman_matches_oracle():
    rng = np.random.default_rng(12)
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    cases = []
    for trial in range(300):
        n = (256, 280, 16)[trial % 3]
        limit = 7 if n == 16 else 15
        f = np.zeros(n, np.uint32)
        k = int(rng.integers(0, n + 1))
        idx = rng.choice(n, k, replace=False)
        kind = trial % 4
        if kind == 0:
            f[idx] = rng.integers(1, 4, k)          # many ties: the heap order decides
        elif kind == 1:
            f[idx] = rng.integers(1, 1 << 20, k)
        elif kind == 2:
            f[idx] = np.array(fib[:k] if k <= 40 else (fib + [1] * (k - 40)), np.uint32)[rng.permutation(k)]  # exceeds the limit
        else:
            f[idx] = (rng.pareto(0.3, k) * 2 + 1).clip(1, 2 ** 27).astype(np.uint32)
        cases.append((f, limit))
    for f, limit in cases:
        ok, lengths, codes = O.build_huffman(f, limit)
        l2 = np.zeros(len(f), np.uint8)
        c2 = np.zeros(len(f), np.uint16)
        ok2 = H.hc_ll_huffman(f.ctypes.data, len(f), limit, l2.ctypes.data, c2.ctypes.data)
        assert bool(ok2) == ok
        assert np.array_equal(l2, lengths) and np.array_equal(c2, codes)
