// zw_quad.cuh -- the luma mode search + final transform of one macroblock with FOUR lanes ("quad") instead of a
// whole warp: every 4x4 block, every candidate, every trellis runs LANE-PRIVATE (one lane = one block / candidate,
// straight-line register code, no shuffles), and a warp walks EIGHT macroblock rows of eight different images at
// once.  The batch supplies the parallelism (32 768 independent rows in 1024 images), so the lanes need not share
// one block.  Round 1's kernels spread one 4x4 block over 8-16 lanes with shuffle butterflies and executed 4.6x
// (pass 2) / 3.2x (pass 1) more thread-instructions than the algorithm has operations.
//
// The lanes of a quad only meet in shared memory (QuadScratch) at __syncwarp(quad mask) points, so quads of one warp
// may diverge freely.  The code is ZW_HD and written against an "executor" X: on the device X::run(f) calls f(lane & 3)
// and syncs the quad; tests/hostcheck runs f(0..3) in turn -- the SAME source is checked on the CPU, macroblock by
// macroblock, against the oracle's P1MB / P2MB dumps before it ever reaches a GPU.
//
// Reference (file:line under /root/reference, src/encoder/vp8.rs unless noted):
//   choose_macroblock_info :2202   pick_best_intra16 :1504   pick_best_intra4 :1790
//   transform_luma_block :2647     transform_luma_blocks_4x4 :2785   trellis_quantize_block cost.rs:788
//   predictors src/common/prediction.rs:164-554
#ifndef ZW_QUAD_CUH
#define ZW_QUAD_CUH
#include "zw_types.cuh"

namespace zw {

constexpr int QG = 4;  // lanes per macroblock row
// The kernel must stay small: a warp streams through the whole macroblock code once per macroblock and an SM has a 32 KB
// instruction cache.  Structural loops are never unrolled and the two big primitives are real (not inlined) functions.
#if defined(__CUDA_ARCH__)
#define ZW_NOUNROLL _Pragma("unroll 1")
#define ZW_NOINLINE __noinline__
#else
#define ZW_NOUNROLL
#define ZW_NOINLINE
#endif
#ifndef ZW_QUAD_PAD
#define ZW_QUAD_PAD 12
#endif
constexpr int QUAD_PAD = ZW_QUAD_PAD;

// Per macroblock row in flight (shared memory on the device): what the four lanes hand to each other.
struct alignas(16) QuadScratch {
  u8 src_y[256];
  u8 yws[17 * 32];     // bordered luma work buffer (prediction.rs LUMA_STRIDE = 32): row 0 / column 0 = borders
  i16 lv[17][16];      // coded levels, zig-zag: [0] Y2, [1..16] Y (the luma part of the record)
  i16 stash[4][16][4]; // block b = 4j + l, natural coefficient k at [j][k][l] (the four lanes of a quad touch neighbouring words):
                       // FDCT coefficients of the I16 mode under evaluation / trellis in-out / I4 winners' levels
  i32 dcs[16];         // DCs of the 16 blocks (Y2 input), then the inverse-WHT output
  i32 tsrc[16];        // TTransform of the source blocks
  u32 part[4][QG];     // per-lane partial sums of an I16 mode: AC cost, SSE, TDisto, non-zero AC levels
  u32 ckey_lo[QG], ckey_hi[QG];  // I4: best candidate key per lane
  u32 csse[QG], crate[QG];
  u32 keys[12];        // I4: prediction-SSE sort keys (sse << 4 | mode)
  u32 y2cost;
  u8 cnz[QG];
  u8 cand_mode[12];
  u8 bmodes[16];
  u8 dtab[48];         // I4: the 23 distinct 3-tap edge filters of the sub-block, DC at [23], TM pixels at [32 + n]
  u8 nzflag[32];       // [0..15] per block, [16..19] top contexts, [20..23] left contexts, [24..27] per-lane flags
  u8 left_y[20];       // [0] corner, [1..16] left column for the next macroblock
  u8 pad_[QUAD_PAD];   // sizeof == 16 (mod 128): the eight quads of a warp sit four banks apart (see QST)
};
static_assert(sizeof(QuadScratch) % 128 == 16, "QuadScratch: keep the quads of a warp four shared-memory banks apart");
#define QST(S, b, k) (S).stash[(b) >> 2][(k)][(b) & 3]

struct QuadConst {
  const u8 (*pidx)[16];  // (mode, pixel) -> index into dtab (TM: 32 + pixel)
  const u16* dtaps;      // slot -> the three edge taps
};

struct QuadMbIn {
  const SegParams* SP;
  CostCtx cc;
  int i4_modes;      // I4 candidates per sub-block: 0 (no I4), 3, 4 or 10
  bool i4_always, trellis;
  int mbx, mby;
  u32 in_top_nz, in_left_nz;
};

struct QuadLumaOut {
  bool use_i4;
  int mode16;
  u32 ynz;         // has_coeffs bit per coded luma block
  int y2nz;        // Y2 has_coeffs (I16 only)
  bool simple_nz;  // any SIMPLE-quantised luma level non-zero (skip test)
};

// get_residual_cost (cost.rs:1670-1729) in flat form: the context of position n is min(|level[n-1]|, 2), known up front,
// so the 16 terms are independent; TYPE / FIRST are compile-time, the band of every position folds to a constant.
ZW_HD constexpr int q_band(int n) { return n < 4 ? n : (n == 4 ? 6 : (n == 5 ? 4 : (n == 6 ? 5 : (n < 15 ? 6 : (n == 15 ? 7 : 0))))); }
template <int TYPE, int FIRST>
ZW_HD u32 q_cost(const i32* lv, int ctx0, const CostCtx& cc) {
  int last = -1, vlast = 0;
#pragma unroll
  for (int i = FIRST; i < 16; i++)
    if (lv[i] != 0) { last = i; vlast = iabs(lv[i]); }
  const u8* pp = cc.probs + TYPE * 264;
  const u32 p0 = pp[(q_band(FIRST) * 3 + ctx0) * 11];
  const u16* lc = cc.level_cost ? cc.level_cost + TYPE * 1632 : nullptr;
  u32 cost = ctx0 == 0 ? bit_cost(1, p0) : 0;
#pragma unroll
  for (int n = FIRST; n < 16; n++) {
    const int v = iabs(lv[n]);
    const int ctx = n == FIRST ? ctx0 : imin(iabs(lv[n - 1]), 2);
    u32 c = ZW_TAB(kLevelFixedCosts)[imin(v, 2047)];
    if (lc) c += lc[(q_band(n) * 3 + ctx) * 68 + imin(v, 67)];
    cost += n <= last ? c : 0u;
  }
  if (last < 15 && last >= 0) cost += bit_cost(0, pp[(ZW_TAB(kEncBands)[last + 1] * 3 + (vlast == 1 ? 1 : 2)) * 11]);
  return last < 0 ? bit_cost(0, p0) : cost;
}

// trellis_quantize_block (cost.rs:788-1006) as a ROLLED loop over the zig-zag positions first..last: the same arithmetic,
// tie rules and outputs as trellis_quantize (zw_cost.cuh, the unrolled form the hostcheck pins against the oracle and
// libwebp's vector), in ~300 instead of ~3400 instructions -- the kernel has to fit an SM's instruction cache -- and
// without the dead positions beyond `last`.  coeffs: natural-order DCT coefficients in, dequantised levels out;
// out: zig-zag levels.  Returns has_nz.
#if defined(__CUDACC__)
__host__ __device__ ZW_NOINLINE
#else
inline
#endif
bool q_trellis(i32* coeffs, i32* out, const Matrix& m, const u16* sharpen, u32 lambda, int first, const CostCtx& cc, int ctype, int ctx0) {
  const i64 MAX_COST = (i64)0x3fffffffffffffffLL;
  const i64 lam = (i64)lambda;
  const i32 q0 = m.q[0], q1 = m.q[1];
  const u32 iq0 = m.iq[0], iq1 = m.iq[1];
  const i32 thresh = (q1 * q1) / 4;
  int last = first - 1;
  bool any = false;  // can any position reach level 1?
  ZW_NOUNROLL
  for (int n = first; n < 16; n++) {
    const int j = ZW_TAB(kZigzag)[n];
    const i32 c = coeffs[j];
    if (c * c > thresh) last = n;
  }
  if (last < 15) last += 1;
  ZW_NOUNROLL
  for (int n = first; n <= last; n++) {
    const int j = ZW_TAB(kZigzag)[n];
    const i32 cs = iabs(coeffs[j]) + (i32)sharpen[j];
    any |= quantdiv((u32)cs, j > 0 ? iq1 : iq0, 1u << 16) >= 1;
  }
  if (!any) {  // a terminal node needs level != 0: nothing can be coded
    ZW_NOUNROLL
    for (int i = first; i < 16; i++) { out[i] = 0; coeffs[i] = 0; }
    return false;
  }
  int best_n = -1, best_delta = 0, best_prev = 0;
  const u8* P = cc.probs + ctype * (8 * 3 * 11);
  const u16* LC = cc.level_cost + ctype * (8 * 3 * 68);
  const int band0 = ZW_TAB(kEncBands)[first];
  i64 best_score = (i64)bit_cost(0, P[(band0 * 3 + ctx0) * 11]) * lam;  // skip: EOB at `first`
  const i64 init_rate = ctx0 == 0 ? (i64)bit_cost(1, P[(band0 * 3 + ctx0) * 11]) : 0;
  i64 sc0 = init_rate * lam, sc1 = sc0;
  const u16* row0 = LC + (band0 * 3 + ctx0) * 68;
  const u16* row1 = row0;
  u32 signs = 0, prev0 = 0, prev1 = 0;  // per-position bit masks
  ZW_NOUNROLL
  for (int n = first; n <= last; n++) {
    const int j = ZW_TAB(kZigzag)[n];
    const i32 q = j > 0 ? q1 : q0;
    const u32 iq = j > 0 ? iq1 : iq0;
    const i32 c = coeffs[j];
    if (c < 0) signs |= 1u << n;
    const i32 cs = iabs(c) + (i32)sharpen[j];
    const i32 level0 = imin(quantdiv((u32)cs, iq, 0), 2047);
    const i32 thresh_level = imin(quantdiv((u32)cs, iq, 1u << 16), 2047);
    const i64 w = ZW_TAB(kWeightTrellis)[j];
    const i64 orig_sq = (i64)(cs * cs);
    const int nband = ZW_TAB(kEncBands)[n + 1];  // kEncBands has 17 entries
    i64 ns0 = MAX_COST, ns1 = MAX_COST;
    const u16 *nrow0 = LC, *nrow1 = LC;
#pragma unroll
    for (int delta = 0; delta < 2; delta++) {
      const i32 level = level0 + delta;
      const int ctx = imin(level, 2);
      const u16* nrow = LC + (nband * 3 + ctx) * 68;  // only read as a predecessor row, i.e. for n + 1 < 16
      if (delta == 0) nrow0 = nrow; else nrow1 = nrow;
      if (level > thresh_level) continue;
      const i32 ne = cs - level * q;
      const i64 base = 256 * (w * ((i64)(ne * ne) - orig_sq));
      const u32 fixed = ZW_TAB(kLevelFixedCosts)[level] + (level > 0 ? 256u : 0u);
      const int lc = imin(level, 67);
      const i64 s0 = sc0 + (i64)(fixed + row0[lc]) * lam;
      const i64 s1 = sc1 + (i64)(fixed + row1[lc]) * lam;
      i64 cur;
      int bp;
      if (s1 < s0) { cur = s1 + base; bp = 1; } else { cur = s0 + base; bp = 0; }
      if (delta == 0) { ns0 = cur; if (bp) prev0 |= 1u << n; } else { ns1 = cur; if (bp) prev1 |= 1u << n; }
      if (level != 0 && cur < best_score) {
        const i64 eob = n < 15 ? (i64)bit_cost(0, P[(nband * 3 + ctx) * 11]) : 0;
        const i64 term = cur + eob * lam;
        if (term < best_score) { best_score = term; best_n = n; best_delta = delta; best_prev = bp; }
      }
    }
    sc0 = ns0; sc1 = ns1; row0 = nrow0; row1 = nrow1;
  }
  // the unwinding needs level0 of every position again: recomputed from the still unmodified coefficients
  i32 lv0[16];
  ZW_NOUNROLL
  for (int n = first; n <= last; n++) {
    const int j = ZW_TAB(kZigzag)[n];
    lv0[n] = imin(quantdiv((u32)(iabs(coeffs[j]) + (i32)sharpen[j]), j > 0 ? iq1 : iq0, 0), 2047);
  }
  ZW_NOUNROLL
  for (int i = first; i < 16; i++) { out[i] = 0; coeffs[i] = 0; }
  if (best_n < 0) return false;
  bool has_nz = false;
  int delta = best_delta;
  ZW_NOUNROLL
  for (int n = best_n; n >= first; n--) {
    const int j = ZW_TAB(kZigzag)[n];
    i32 level = lv0[n] + delta;
    if ((signs >> n) & 1) level = -level;
    out[n] = level;
    has_nz |= level != 0;
    coeffs[j] = level * (j > 0 ? q1 : q0);
    delta = (n == best_n) ? best_prev : (int)(((delta ? prev1 : prev0) >> n) & 1);
  }
  return has_nz;
}

// zig-zag position of natural coefficient index k (the inverse of kZigzag), nibble k of the constant
ZW_HD int q_zinv(int k) { return (int)((0xFEA9DB83C7426510ull >> (4 * k)) & 15); }

// a lane's best I4 candidate so far: reconstruction (4 packed rows) and natural-order levels (8 packed pairs)
struct CandBest {
  u64 key;
  u32 rec[4], lv[8];
  int mode;
};

// one value per lane of the quad: a register on the device (LANES = 1), an array on the host (LANES = QG)
template <class T, int LANES>
struct LaneVar {
  T v[LANES];
  ZW_HD T& operator()(int q) { return v[LANES == 1 ? 0 : q]; }
};

// 16x16 whole-block predictors for one 4x4 block (bx, by): one formula for the four modes, clip255(T'[x] + L'[y] + base)
// with T' = T for V / TM (else 0), L' = L for H / TM (else 0), base = dc for DC, -P for TM (prediction.rs:164-324).
ZW_HD void q_pred_block(const u8* ws, int mode, int bx, int by, int dcv, i32* pr) {
  const bool useT = (mode & 1) != 0, useL = mode >= 2;
  i32 T[4], L[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    T[k] = useT ? (i32)ws[1 + bx * 4 + k] : 0;
    L[k] = useL ? (i32)ws[(1 + by * 4 + k) * 32] : 0;
  }
  const i32 base = mode == 0 ? dcv : (mode == 3 ? -(i32)ws[0] : 0);
#pragma unroll
  for (int k = 0; k < 16; k++) pr[k] = clip255(T[k & 3] + L[k >> 2] + base);
}

ZW_HD void q_load_src(const u8* src_y, int bx, int by, i32* px) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const u32 w = *reinterpret_cast<const u32*>(&src_y[(by * 4 + r) * 16 + bx * 4]);
    px[4 * r] = (i32)(w & 255u); px[4 * r + 1] = (i32)((w >> 8) & 255u); px[4 * r + 2] = (i32)((w >> 16) & 255u); px[4 * r + 3] = (i32)(w >> 24);
  }
}

// Complexity left behind by a macroblock for the row below (out_top) and the MB to the right (out_left):
// encode_residual_data / record_residual_stats bookkeeping + the skip clear (vp8.rs:1367-1374, :1468-1474, Complexity::clear :138).
ZW_HD void q_complexity_after(bool is_b, bool skip, int y2nz, u32 ynz, u32 uvnz, u32 in_top, u32 in_left, u32& out_top, u32& out_left) {
  u32 yb = ynz, ub = uvnz & 15, vb = (uvnz >> 4) & 15;
  int y2t = is_b ? (int)(in_top & 1) : (y2nz != 0);
  int y2l = is_b ? (int)(in_left & 1) : (y2nz != 0);
  if (skip) {
    yb = 0; ub = 0; vb = 0;
    if (!is_b) { y2t = 0; y2l = 0; }
  }
  out_top = (u32)y2t | (((yb >> 12) & 15) << 1) | (((ub >> 2) & 3) << 5) | (((vb >> 2) & 3) << 7);
  const u32 yl = ((yb >> 3) & 1) | (((yb >> 7) & 1) << 1) | (((yb >> 11) & 1) << 2) | (((yb >> 15) & 1) << 3);
  const u32 ul = ((ub >> 1) & 1) | (((ub >> 3) & 1) << 1);
  const u32 vl = ((vb >> 1) & 1) | (((vb >> 3) & 1) << 1);
  out_left = (u32)y2l | (yl << 1) | (ul << 5) | (vl << 7);
}

// What the three phases of a macroblock hand to each other (identical in the four lanes of a quad).
struct QuadMbState {
  int dc16, best16_mode;
  u64 i16_score;
  bool use_i4;
};

// The luma decision of one macroblock in three phases (the kernel puts a CTA barrier between them so that the warps of an
// SM run the same code at the same time: the three loop bodies each fit the instruction cache, all of it does not).
// Expects S.src_y, the borders of S.yws and S.lv zeroed; leaves the reconstruction in S.yws, the coded zig-zag levels in
// S.lv and the sub-block modes in S.bmodes.
// ---- phase 1: pick_best_intra16 (vp8.rs:1504-1681); lane q owns blocks q, q + 4, q + 8, q + 12 of every mode ----
template <int LANES, class X>
ZW_HD void quad_i16(X& x, QuadScratch& S, const QuadConst& K, const QuadMbIn& in, QuadMbState& st) {
  (void)K;
  const SegParams SP = *in.SP;  // a private copy: the quantiser entries stay in registers across the shared-memory traffic
  const CostCtx& cc = in.cc;
  const int mbx = in.mbx, mby = in.mby;
  const i64 I64MAX = 0x7fffffffffffffffLL;

  // ===== pick_best_intra16 (vp8.rs:1504-1681): lane q owns blocks 4q .. 4q+3 of every mode =====
  int dc16;
  {
    int s = 0;
    for (int k = 0; k < 16; k++) s += (mby != 0 ? S.yws[1 + k] : 0) + (mbx != 0 ? S.yws[(1 + k) * 32] : 0);
    const int shf = 3 + (mbx != 0) + (mby != 0);
    dc16 = (mbx == 0 && mby == 0) ? 128 : (s + (1 << (shf - 1))) >> shf;  // predict_dcpred :183
  }
  x.run([&](int q) {  // TTransform of the source blocks (cost.rs:73) + is_flat_source_16 (cost.rs:177)
    const u32 v0 = S.src_y[0] * 0x01010101u;
    bool same = true;
    ZW_NOUNROLL
    for (int j = 0; j < 4; j++) {
      const int b = 4 * j + q;
      i32 px[16];
      q_load_src(S.src_y, b & 3, b >> 2, px);
      S.tsrc[b] = t_transform16(px, ZW_TAB(kWeightY));
#pragma unroll
      for (int r = 0; r < 4; r++) same &= *reinterpret_cast<const u32*>(&S.src_y[((b >> 2) * 4 + r) * 16 + (b & 3) * 4]) == v0;
    }
    S.nzflag[24 + q] = same;
  });
  const bool is_flat = S.nzflag[24] && S.nzflag[25] && S.nzflag[26] && S.nzflag[27];
  i64 best16_score = I64MAX;
  int best16_mode = 0;
  u32 best16_cc = 0, best16_mc = 0, best16_d = 0;
  i32 best16_sd = 0;
  ZW_NOUNROLL
  for (int mode = 0; mode < 4; mode++) {  // 0 DC, 1 V, 2 H, 3 TM (MODES order, vp8.rs:1509)
    const bool avail = !((mode == 1 && mby == 0) || (mode == 2 && mbx == 0) || (mode == 3 && (mbx == 0 || mby == 0)));
    if (!avail) continue;
    x.run([&](int q) {
      ZW_NOUNROLL
    for (int j = 0; j < 4; j++) {
        const int b = 4 * j + q, bx = b & 3, by = b >> 2;
        i32 c[16], pr[16];
        q_pred_block(S.yws, mode, bx, by, dc16, pr);
        q_load_src(S.src_y, bx, by, c);
#pragma unroll
        for (int k = 0; k < 16; k++) c[k] -= pr[k];
        fdct4x4(c);
        S.dcs[b] = c[0];
#pragma unroll
        for (int k = 1; k < 16; k++) QST(S, b, k) = (i16)c[k];
      }
    });
    x.run([&](int q) {  // Y2: the 16 DCs (raster block order = coefficient position), one lane
      if (q != 0) return;
      i32 y2[16], lvq[16];
#pragma unroll
      for (int k = 0; k < 16; k++) y2[k] = S.dcs[k];
      wht4x4(y2);
#pragma unroll
      for (int k = 0; k < 16; k++) { lvq[k] = quantize_coeff(y2[k], SP.y2, k); y2[k] = dequantize(lvq[k], SP.y2, k); }
      S.y2cost = q_cost<1, 0>(lvq, 0, cc);
      iwht4x4(y2);
#pragma unroll
      for (int k = 0; k < 16; k++) S.dcs[k] = y2[k];
    });
    x.run([&](int q) {
      u32 cost_ac = 0, sse = 0, td = 0, nzc = 0;
      ZW_NOUNROLL
    for (int j = 0; j < 4; j++) {
        const int b = 4 * j + q, bx = b & 3, by = b >> 2;
        i32 c[16], pr[16], lvq[16], src[16];
        q_pred_block(S.yws, mode, bx, by, dc16, pr);
        lvq[0] = 0;
#pragma unroll
        for (int k = 1; k < 16; k++) { lvq[k] = quantize_coeff((i32)QST(S, b, k), SP.y1, k); nzc += lvq[k] != 0; }
        cost_ac += q_cost<0, 1>(lvq, 0, cc);
#pragma unroll
        for (int k = 1; k < 16; k++) c[k] = dequantize(lvq[k], SP.y1, k);
        c[0] = S.dcs[b];
        idct4x4(c);
        q_load_src(S.src_y, bx, by, src);
#pragma unroll
        for (int k = 0; k < 16; k++) {
          c[k] = clip255(pr[k] + c[k]);
          const i32 df = src[k] - c[k];
          sse += (u32)(df * df);
        }
        td += (u32)(iabs(t_transform16(c, ZW_TAB(kWeightY)) - S.tsrc[b]) >> 5);  // tdisto_4x4 (cost.rs:122)
      }
      S.part[0][q] = cost_ac; S.part[1][q] = sse; S.part[2][q] = td; S.part[3][q] = nzc;
    });
    {
      u32 sum[4];
      for (int k = 0; k < 4; k++) sum[k] = S.part[k][0] + S.part[k][1] + S.part[k][2] + S.part[k][3];
      const u32 coeff_cost = S.y2cost + sum[0];
      i32 sd = SP.tlambda > 0 ? (i32)(((i32)SP.tlambda * (i32)sum[2] + 128) >> 8) : 0;
      u32 dfin = sum[1];
      if (is_flat && sum[3] == 0) { dfin = dfin * 2; sd = sd * 2; }  // is_flat_coeffs(.., 16, 0)
      const u32 mode_cost = ZW_TAB(kFixedCostsI16)[mode];
      const i64 score = ((i64)mode_cost + (i64)coeff_cost) * (i64)SP.lambda_i16 + 256 * ((i64)dfin + (i64)sd);
      if (score < best16_score) {  // strict <: the first of equal scores stays
        best16_score = score; best16_mode = mode; best16_cc = coeff_cost; best16_mc = mode_cost; best16_d = dfin; best16_sd = sd;
      }
    }
  }
  {
    const i64 fs = ((i64)best16_mc + (i64)best16_cc) * (i64)SP.lambda_mode + 256 * ((i64)best16_d + (i64)best16_sd);
    st.i16_score = (u64)(fs > 0 ? fs : 0);
  }
  st.dc16 = dc16; st.best16_mode = best16_mode;
  // gated as in choose_macroblock_info (:2210-2231)
  st.use_i4 = in.i4_modes > 0 && (in.i4_always || st.i16_score > 211ull * (u64)SP.lambda_mode || best16_mode != 0);
}

// ---- phase 2: pick_best_intra4 (vp8.rs:1790-2036) ----
template <int LANES, class X>
ZW_HD void quad_i4(X& x, QuadScratch& S, const QuadConst& K, const QuadMbIn& in, QuadMbState& st) {
  const SegParams SP = *in.SP;
  const CostCtx& cc = in.cc;
  const u64 i16_score = st.i16_score;
  bool use_i4 = st.use_i4;
  if (use_i4) {
    const int max_modes = in.i4_modes;
    const u32 lam_i4 = SP.lambda_i4, lam_mode = SP.lambda_mode;
    u64 running = 211ull * (u64)lam_mode;
    u32 total_mode_cost = 0, tnz4 = 0, lnz4 = 0;  // MB-local non-zero context bits
    LaneVar<CandBest, LANES> cb;
    ZW_NOUNROLL
    for (int i = 0; i < 16 && use_i4; i++) {
      const int sbx = i & 3, sby = i >> 2, x0 = 1 + 4 * sbx, y0 = 1 + 4 * sby;
      const int top_ctx = sby == 0 ? 0 : S.bmodes[i - 4];
      const int left_ctx = sbx == 0 ? 0 : S.bmodes[i - 1];
      const int ctx0 = (sby == 0 ? 0 : (int)((tnz4 >> sbx) & 1)) + (sbx == 0 ? 0 : (int)((lnz4 >> sby) & 1));
      x.run([&](int q) {  // the 23 distinct 3-tap filters of the 13 edge pixels, DC, TM (each lane a share)
        u8 e[13];
#pragma unroll
        for (int k = 0; k < 4; k++) e[k] = S.yws[(y0 + 3 - k) * 32 + x0 - 1];
#pragma unroll
        for (int k = 4; k < 13; k++) e[k] = S.yws[(y0 - 1) * 32 + x0 - 5 + k];
        {
          // the 23 distinct 3-tap values (ZW_DTAPS_INIT): 0..10 avg3(e[k], e[k+1], e[k+2]); 11 avg3(e11, e12, e12); 12 avg3(e0, e0, e1);
          // 13..21 avg2(e[k-13], e[k-12]); 22 = e0.  Static indexing: every lane evaluates all, and stores its share.
          i32 v[23];
#pragma unroll
          for (int k = 0; k < 11; k++) v[k] = ((i32)e[k] + 2 * (i32)e[k + 1] + (i32)e[k + 2] + 2) >> 2;
          v[11] = ((i32)e[11] + 3 * (i32)e[12] + 2) >> 2;
          v[12] = (3 * (i32)e[0] + (i32)e[1] + 2) >> 2;
#pragma unroll
          for (int k = 13; k < 22; k++) v[k] = ((i32)e[k - 13] + (i32)e[k - 12] + 1) >> 1;
          v[22] = (i32)e[0];
#pragma unroll
          for (int k = 0; k < 23; k++) if ((k & 3) == q) S.dtab[k] = (u8)v[k];
        }
        if (q == 3) S.dtab[23] = (u8)((4 + e[5] + e[6] + e[7] + e[8] + e[0] + e[1] + e[2] + e[3]) >> 3);
#pragma unroll
        for (int k = 0; k < 4; k++) S.dtab[32 + 4 * q + k] = (u8)clip255((i32)e[3 - q] + (i32)e[5 + k] - (i32)e[4]);  // TM row q
      });
      x.run([&](int q) {  // prediction SSE of the ten modes (lane q: modes q, q + 4, q + 8)
        i32 src[16];
        q_load_src(S.src_y, sbx, sby, src);
        ZW_NOUNROLL
        for (int m = q; m < 10; m += 4) {
          u32 sse = 0;
#pragma unroll
          for (int k = 0; k < 16; k++) { const i32 df = src[k] - (i32)S.dtab[K.pidx[m][k]]; sse += (u32)(df * df); }
          S.keys[m] = (sse << 4) | (u32)m;
        }
      });
      x.run([&](int q) {  // ascending key order == stable ascending SSE order (an insertion sort at this length, Q11)
        u32 kk[10];
#pragma unroll
        for (int j = 0; j < 10; j++) kk[j] = S.keys[j];
        ZW_NOUNROLL
        for (int m = q; m < 10; m += 4) {
          const u32 mine = S.keys[m];
          int rank = 0;
#pragma unroll
          for (int j = 0; j < 10; j++) rank += kk[j] < mine;
          if (rank < max_modes) S.cand_mode[rank] = (u8)m;
        }
      });
      x.run([&](int q) {  // full RD of the best `max_modes` candidates, one per lane and round
        i32 src[16];
        q_load_src(S.src_y, sbx, sby, src);
        CandBest& B = cb(q);
        B.key = ~0ull;
        ZW_NOUNROLL
        for (int rank = q; rank < max_modes; rank += 4) {
          const int m = S.cand_mode[rank];
          i32 c[16], pr[16], lvq[16];
#pragma unroll
          for (int k = 0; k < 16; k++) { pr[k] = (i32)S.dtab[K.pidx[m][k]]; c[k] = src[k] - pr[k]; }
          fdct4x4(c);
          u32 nz = 0;
#pragma unroll
          for (int k = 0; k < 16; k++) { lvq[k] = quantize_coeff(c[k], SP.y1, k); nz |= lvq[k] != 0; }
          const u32 coeff_cost = q_cost<3, 0>(lvq, ctx0, cc);
#pragma unroll
          for (int k = 0; k < 16; k++) c[k] = dequantize(lvq[k], SP.y1, k);
          idct4x4(c);
          u32 sse = 0;
#pragma unroll
          for (int k = 0; k < 16; k++) { c[k] = clip255(pr[k] + c[k]); const i32 df = src[k] - c[k]; sse += (u32)(df * df); }
          const u32 rate = ZW_TAB(kFixedCostsI4)[(top_ctx * 10 + left_ctx) * 10 + m] + coeff_cost;
          const u64 score = (u64)sse * 256ull + (u64)(rate & 0xffffu) * (u64)lam_i4;  // u16 truncation (Q8)
          const u64 key = (score << 4) | (u64)rank;  // the rank makes keys unique: min == first best
          if (key < B.key) {
            B.key = key; B.mode = m;
#pragma unroll
            for (int k = 0; k < 4; k++) B.rec[k] = (u32)c[4 * k] | ((u32)c[4 * k + 1] << 8) | ((u32)c[4 * k + 2] << 16) | ((u32)c[4 * k + 3] << 24);
#pragma unroll
            for (int k = 0; k < 8; k++) B.lv[k] = ((u32)lvq[2 * k] & 0xffffu) | ((u32)lvq[2 * k + 1] << 16);
            S.csse[q] = sse; S.crate[q] = rate; S.cnz[q] = (u8)nz;
          }
        }
        S.ckey_lo[q] = (u32)B.key; S.ckey_hi[q] = (u32)(B.key >> 32);
      });
      int w = 0;  // the winning lane (keys are unique)
      for (int k = 1; k < QG; k++)
        if ((((u64)S.ckey_hi[k] << 32) | S.ckey_lo[k]) < (((u64)S.ckey_hi[w] << 32) | S.ckey_lo[w])) w = k;
      const u32 best_sse = S.csse[w], best_rate = S.crate[w], best_nz = S.cnz[w];
      x.run([&](int q) {
        if (q != w) return;
        const CandBest& B = cb(q);
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
          for (int k = 0; k < 4; k++) S.yws[(y0 + r) * 32 + x0 + k] = (u8)(B.rec[r] >> (8 * k));
        }
#pragma unroll
        for (int k = 0; k < 8; k++) { QST(S, i, 2 * k) = (i16)(B.lv[k] & 0xffffu); QST(S, i, 2 * k + 1) = (i16)(B.lv[k] >> 16); }
        S.bmodes[i] = (u8)B.mode;
      });
      const int wmode = S.bmodes[i];
      tnz4 = (tnz4 & ~(1u << sbx)) | (best_nz << sbx);
      lnz4 = (lnz4 & ~(1u << sby)) | (best_nz << sby);
      total_mode_cost += ZW_TAB(kFixedCostsI4)[(top_ctx * 10 + left_ctx) * 10 + wmode];
      running += (u64)best_sse * 256ull + (u64)(best_rate & 0xffffu) * (u64)lam_mode;
      if (running >= i16_score || total_mode_cost > 16384u) use_i4 = false;
    }
  }

  st.use_i4 = use_i4;
}

// ---- phase 3: final luma transform -> coded levels + reconstruction ----
template <int LANES, class X>
ZW_HD QuadLumaOut quad_final(X& x, QuadScratch& S, const QuadConst& K, const QuadMbIn& in, const QuadMbState& st) {
  const SegParams SP = *in.SP;
  const CostCtx& cc = in.cc;
  const bool use_i4 = st.use_i4;
  const int best16_mode = st.best16_mode, dc16 = st.dc16;
  bool any_simple_nz = false;
  u32 ynz = 0;
  int y2nz = 0;
  if (!use_i4) {
    // ---- transform_luma_block (vp8.rs:2647-2780) ----
    x.run([&](int q) {
      ZW_NOUNROLL
    for (int j = 0; j < 4; j++) {
        const int b = 4 * j + q, bx = b & 3, by = b >> 2;
        i32 c[16], pr[16];
        q_pred_block(S.yws, best16_mode, bx, by, dc16, pr);
        q_load_src(S.src_y, bx, by, c);
#pragma unroll
        for (int k = 0; k < 16; k++) c[k] -= pr[k];
        fdct4x4(c);
        S.dcs[b] = c[0];
#pragma unroll
        for (int k = 1; k < 16; k++) QST(S, b, k) = (i16)c[k];
      }
      if (q == 0 && in.trellis) {
#pragma unroll
        for (int k = 0; k < 4; k++) { S.nzflag[16 + k] = (in.in_top_nz >> (1 + k)) & 1; S.nzflag[20 + k] = (in.in_left_nz >> (1 + k)) & 1; }
      }
    });
    x.run([&](int q) {  // Y2 (never trellis-quantised, vp8.rs:701)
      if (q != 0) return;
      i32 y2[16];
      u32 nz = 0;
#pragma unroll
      for (int k = 0; k < 16; k++) y2[k] = S.dcs[k];
      wht4x4(y2);
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const i32 l = quantize_coeff(y2[k], SP.y2, k);
        nz |= l != 0;
        y2[k] = dequantize(l, SP.y2, k);
        S.lv[0][q_zinv(k)] = (i16)l;
      }
      S.nzflag[24] = (u8)nz;
      iwht4x4(y2);
#pragma unroll
      for (int k = 0; k < 16; k++) S.dcs[k] = y2[k];
    });
    y2nz = S.nzflag[24] != 0;
    x.run([&](int q) {  // simple quantisation: the skip test always (Q13), the coded levels without trellis
      u32 simple = 0;
      ZW_NOUNROLL
    for (int j = 0; j < 4; j++) {
        const int b = 4 * j + q;
        i32 lvq[16];
        lvq[0] = 0;
        u32 nz = 0;
#pragma unroll
        for (int k = 1; k < 16; k++) { lvq[k] = quantize_coeff((i32)QST(S, b, k), SP.y1, k); nz |= lvq[k] != 0; }
        simple |= nz;
        if (!in.trellis) {
#pragma unroll
          for (int k = 0; k < 16; k++) S.lv[1 + b][q_zinv(k)] = (i16)lvq[k];
#pragma unroll
          for (int k = 1; k < 16; k++) QST(S, b, k) = (i16)dequantize(lvq[k], SP.y1, k);
          S.nzflag[b] = (u8)nz;
        }
      }
      S.nzflag[28 + q] = (u8)simple;
    });
    any_simple_nz = y2nz != 0 || S.nzflag[28] || S.nzflag[29] || S.nzflag[30] || S.nzflag[31];
    if (in.trellis) {
      // trellis with the nz context chained in raster order (:2685-2728): blocks on one anti-diagonal are independent
      ZW_NOUNROLL
      for (int d = 0; d < 7; d++) {
        x.run([&](int q) {
          const int bx = d < 4 ? d - q : 3 - q, by = d < 4 ? q : d - 3 + q;  // q-th block of diagonal d
          if (bx < 0 || by > 3) return;
          const int b = by * 4 + bx;
          const int ctx0 = imin((int)S.nzflag[20 + by] + (int)S.nzflag[16 + bx], 2);
          i32 c[16], zz[16];
          c[0] = 0;
#pragma unroll
          for (int k = 1; k < 16; k++) c[k] = (i32)QST(S, b, k);
#pragma unroll
          for (int k = 0; k < 16; k++) zz[k] = 0;
          const bool nz = q_trellis(c, zz, in.SP->y1, in.SP->sharpen, SP.lambda_trellis_i16, 1, cc, 0, ctx0);
#pragma unroll
          for (int k = 1; k < 16; k++) { QST(S, b, k) = (i16)c[k]; S.lv[1 + b][k] = (i16)zz[k]; }
          S.lv[1 + b][0] = 0;
          S.nzflag[b] = nz;
          S.nzflag[24 + q] = nz;  // contexts are updated after the whole diagonal has read them
        });
        x.run([&](int q) {
          const int bx = d < 4 ? d - q : 3 - q, by = d < 4 ? q : d - 3 + q;
          if (bx < 0 || by > 3) return;
          S.nzflag[16 + bx] = S.nzflag[24 + q];
          S.nzflag[20 + by] = S.nzflag[24 + q];
        });
      }
    }
    x.run([&](int q) {  // reconstruction
      ZW_NOUNROLL
    for (int j = 0; j < 4; j++) {
        const int b = 4 * j + q, bx = b & 3, by = b >> 2;
        i32 c[16], pr[16];
        q_pred_block(S.yws, best16_mode, bx, by, dc16, pr);
        c[0] = S.dcs[b];
#pragma unroll
        for (int k = 1; k < 16; k++) c[k] = (i32)QST(S, b, k);
        idct4x4(c);
        // (the prediction only reads row 0 / column 0 of yws, which stay untouched)
#pragma unroll
        for (int k = 0; k < 16; k++) S.yws[(1 + by * 4 + (k >> 2)) * 32 + 1 + bx * 4 + (k & 3)] = (u8)clip255(pr[k] + c[k]);
      }
    });
    for (int b = 0; b < 16; b++) ynz |= (u32)(S.nzflag[b] != 0) << b;
  } else if (!in.trellis) {
    // ---- transform_luma_blocks_4x4 without trellis == what the search already produced ----
    x.run([&](int q) {
      ZW_NOUNROLL
    for (int j = 0; j < 4; j++) {
        const int b = 4 * j + q;
        u32 nz = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) { const i16 v = QST(S, b, k); S.lv[1 + b][q_zinv(k)] = v; nz |= v != 0; }
        S.nzflag[b] = (u8)nz;
      }
      if (q == 0) {
#pragma unroll
        for (int k = 0; k < 16; k++) S.lv[0][k] = 0;
      }
    });
    for (int b = 0; b < 16; b++) ynz |= (u32)(S.nzflag[b] != 0) << b;
    any_simple_nz = ynz != 0;
  } else {
    // ---- transform_luma_blocks_4x4 with trellis (vp8.rs:2785-2916).  The reference walks the 16 sub-blocks in raster
    //      order; block (x, y) only needs its left, top and top-right neighbours (prediction, non-zero contexts), so
    //      blocks with equal x + 2y are independent: ten rounds, lanes 0 and 1 taking one block each in six of them ----
    u32 tnz = (in.in_top_nz >> 1) & 15, lnz = (in.in_left_nz >> 1) & 15;
    x.run([&](int q) {
      if (q == 0) {
#pragma unroll
        for (int k = 0; k < 16; k++) S.lv[0][k] = 0;
      }
      S.nzflag[28 + q] = 0;
    });
    ZW_NOUNROLL
    for (int rd = 0; rd < 10; rd++) {
      // round -> block of lane 0 / lane 1 (-1: none): {0,-} {1,-} {2,4} {3,5} {6,8} {7,9} {10,12} {11,13} {14,-} {15,-}
      const int ba = rd < 4 ? rd : (rd < 8 ? 6 + ((rd - 4) >> 1) * 4 + (rd & 1) : 6 + rd);
      const int bb = (rd >= 2 && rd < 8) ? ba + 2 : -1;
      x.run([&](int q) {
        const int i = q == 0 ? ba : (q == 1 ? bb : -1);
        if (i < 0) return;
        const int sbx = i & 3, sby = i >> 2, x0 = 1 + 4 * sbx, y0 = 1 + 4 * sby;
        const int bmode = S.bmodes[i];
        u8 e[13];
#pragma unroll
        for (int k = 0; k < 4; k++) e[k] = S.yws[(y0 + 3 - k) * 32 + x0 - 1];
#pragma unroll
        for (int k = 4; k < 13; k++) e[k] = S.yws[(y0 - 1) * 32 + x0 - 5 + k];
        i32 c[16], pr[16], zz[16];
        q_load_src(S.src_y, sbx, sby, c);
#pragma unroll
        for (int k = 0; k < 16; k++) { pr[k] = predict4_pixel_lut(e, bmode, k, K.dtaps, K.pidx); c[k] -= pr[k]; }
        fdct4x4(c);
        u32 simple = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) { simple |= quantize_coeff(c[k], SP.y1, k) != 0; zz[k] = 0; }
        const int ctx0 = imin((int)((lnz >> sby) & 1) + (int)((tnz >> sbx) & 1), 2);
        const bool nz = q_trellis(c, zz, in.SP->y1, in.SP->sharpen, SP.lambda_trellis_i4, 0, cc, 3, ctx0);
#pragma unroll
        for (int k = 0; k < 16; k++) S.lv[1 + i][k] = (i16)zz[k];
        idct4x4(c);
#pragma unroll
        for (int k = 0; k < 16; k++) S.yws[(y0 + (k >> 2)) * 32 + x0 + (k & 3)] = (u8)clip255(pr[k] + c[k]);
        S.nzflag[i] = nz;
        if (simple) S.nzflag[28 + q] = 1;
      });
      {
        const u32 nza = S.nzflag[ba] != 0;
        const int ax = ba & 3, ay = ba >> 2;
        tnz = (tnz & ~(1u << ax)) | (nza << ax);
        lnz = (lnz & ~(1u << ay)) | (nza << ay);
        ynz |= nza << ba;
        if (bb >= 0) {
          const u32 nzb = S.nzflag[bb] != 0;
          const int bx2 = bb & 3, by2 = bb >> 2;
          tnz = (tnz & ~(1u << bx2)) | (nzb << bx2);
          lnz = (lnz & ~(1u << by2)) | (nzb << by2);
          ynz |= nzb << bb;
        }
      }
    }
    any_simple_nz = S.nzflag[28] || S.nzflag[29];
  }
  QuadLumaOut R;
  R.use_i4 = use_i4; R.mode16 = best16_mode; R.ynz = ynz; R.y2nz = y2nz; R.simple_nz = any_simple_nz;
  return R;
}

template <int LANES, class X>
ZW_HD QuadLumaOut quad_luma_mb(X& x, QuadScratch& S, const QuadConst& K, const QuadMbIn& in) {
  QuadMbState st;
  quad_i16<LANES>(x, S, K, in, st);
  quad_i4<LANES>(x, S, K, in, st);
  return quad_final<LANES>(x, S, K, in, st);
}

}  // namespace zw
#endif
