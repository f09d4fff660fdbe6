"""CPU restatement of two lane mappings of the luma wavefront kernel (image_webp_b200/csrc/zw_search.cuh, luma_mb):
the I4 prediction-SSE ranking on packed bytes (four lanes per mode, one pixel row each; modes 8 and 9 in a second step;
key owners lanes 4m and lanes 1 / 5) and the merge of the four candidate lane groups (a 64-bit minimum as two 32-bit
minima).  The kernel itself is pinned byte for byte by the GPU parity tests; this pins the index algebra against the
plain form of the reference (vp8.rs:1897-1920: SSE of the ten 4x4 predictors, stable ascending order) on the CPU."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tables():
    txt = open(os.path.join(ROOT, "image_webp_b200", "csrc", "zw_prims.cuh")).read()
    body = txt[txt.index("#define ZW_PRED_IDX_INIT"):]
    body = body[:body.index("// Host-checkable")]
    rows = re.findall(r"\{([0-9,\s]+)\}", body)
    pidx = np.array([[int(v) for v in r.split(",")] for r in rows], dtype=np.int64)
    assert pidx.shape == (10, 16)
    pidx[1] = 32 + np.arange(16)  # the kernel's prologue: the TM pixels live in dtab[32 + n]
    taps = txt[txt.index("#define ZW_DTAPS_INIT"):]
    taps = [int(v) for v in re.findall(r"\d+", taps[:taps.index("#define ZW_PRED_IDX_INIT")].split("{", 1)[1])][:32]
    return pidx, taps


def _dtab(e, taps):
    """e[0..12] = L3 L2 L1 L0 P A0..A7 -> the 48-byte predictor table of one sub-block (pred4_prepare)."""
    d = np.zeros(48, dtype=np.int64)
    for k in range(23):
        a, b, c = taps[k] & 15, (taps[k] >> 4) & 15, (taps[k] >> 8) & 15
        d[k] = (e[a] + 2 * e[b] + e[c] + 2) >> 2
    d[23] = (sum(e[0:4]) + sum(e[5:9]) + 4) >> 3
    for n in range(16):
        d[32 + n] = min(255, max(0, e[3 - (n >> 2)] + e[5 + (n & 3)] - e[4]))
    return d


def test_packed_sse_ranking_matches_the_plain_form():
    pidx, taps = _tables()
    rng = np.random.default_rng(7)
    for trial in range(400):
        flat = trial % 5 == 0  # flat edges: many equal SSEs, so the tie order (mode index) is exercised
        e = np.full(13, int(rng.integers(0, 256))) if flat else rng.integers(0, 256, 13)
        src = rng.integers(0, 256, 16) if trial % 3 else np.full(16, int(rng.integers(0, 256)))
        d = _dtab(e, taps)
        plain = [int(((src - d[pidx[m]]) ** 2).sum()) for m in range(10)]
        want = [m for _, m in sorted((s, m) for m, s in enumerate(plain))]
        # the kernel's lanes
        sA, sB = np.zeros(32, dtype=np.int64), np.zeros(32, dtype=np.int64)
        for lane in range(32):
            row = lane & 3
            for acc, mode in ((sA, lane >> 2), (sB, 8 + ((lane >> 2) & 1))):
                acc[lane] = int(((src[4 * row:4 * row + 4] - d[pidx[mode][4 * row:4 * row + 4]]) ** 2).sum())
        for x in (1, 2):  # two xor-shuffles finish a mode
            sA = sA + sA[np.arange(32) ^ x]
            sB = sB + sB[np.arange(32) ^ x]
        keys = []
        for lane in range(32):
            if lane & 3 == 0:
                keys.append((int(sA[lane]) << 4) | (lane >> 2))
            elif lane in (1, 5):
                keys.append((int(sB[lane]) << 4) | (8 + (lane >> 2)))
            else:
                keys.append(0xFFFFFFFF)
        assert max(plain) << 4 < 0xFFFFFFFF
        got = []
        for _ in range(10):
            kmin = min(keys)
            got.append(kmin & 15)
            keys[keys.index(kmin)] = 0xFFFFFFFF
        assert got == want


def test_group_merge_by_two_32_bit_minima():
    rng = np.random.default_rng(11)
    for trial in range(2000):
        n_cand = int(rng.integers(1, 11))           # method 4: four candidates, methods 5/6: ten (three steps)
        his = rng.integers(0, 3, 10) if trial % 2 else rng.integers(0, 1 << 20, 10)  # force equal high words often
        group_best = [None] * 4
        for rank in range(n_cand):                  # rank r is evaluated by lane group r & 3
            key = (int(his[rank]) << 32) | (int(rng.integers(0, 1 << 28)) << 4) | rank
            g = rank & 3
            if group_best[g] is None or key < group_best[g]:
                group_best[g] = key
        lanes = [group_best[l >> 3] if group_best[l >> 3] is not None else (1 << 64) - 1 for l in range(32)]
        mhi = min(k >> 32 for k in lanes)
        mlo = min((k & 0xFFFFFFFF) if (k >> 32) == mhi else 0xFFFFFFFF for k in lanes)
        merged = (mhi << 32) | mlo
        assert merged == min(lanes)
        wsrc = (mlo & 3) << 3                       # first lane of the winning group
        assert lanes[wsrc] == merged
