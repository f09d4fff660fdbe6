// zw_cost.cuh -- rate model, trellis quantiser and token statistics, per lane (sm_100a).
//
// Reference semantics (file:line under /root/reference):
//   vp8_bit_cost            src/encoder/cost.rs:40
//   LevelCosts::calculate   src/encoder/cost.rs:1500-1546   (level_cost_build)
//   get_residual_cost       src/encoder/cost.rs:1670-1729   (residual_cost)
//   trellis_quantize_block  src/encoder/cost.rs:788-1006    (trellis_quantize)
//   record_coeffs           src/encoder/cost.rs:1297-1397   (token_events)
// Quirks kept: Q1 pass-1 tables are all zero; Q7 costs see levels in
// natural order; Q10 skip_eob never cleared in the statistics; Q14 trellis details.
#ifndef ZW_COST_CUH
#define ZW_COST_CUH
#include "zw_prims.cuh"

namespace zw {

namespace host {
#define ZW_TABLE_QUAL static const
#include "zw_tables.inc"
#undef ZW_TABLE_QUAL
}  // namespace host
#if defined(__CUDACC__)
namespace dev {
#define ZW_TABLE_QUAL __device__ const
#include "zw_tables.inc"
#undef ZW_TABLE_QUAL
}  // namespace dev
#endif

#if defined(__CUDA_ARCH__)
#define ZW_TAB(x) ::zw::dev::x
#else
#define ZW_TAB(x) ::zw::host::x
#endif

ZW_HD u32 bit_cost(int bit, u32 prob) { return bit ? ZW_TAB(kEntropyCost)[255 - prob] : ZW_TAB(kEntropyCost)[prob]; }

// Rate-model view of one image for one pass.
struct CostCtx {
  const u8* probs;        // [4][8][3][11] token probabilities the estimates read (p0 terms)
  const u16* level_cost;  // [4][8][3][68] variable level costs; nullptr == the all-zero pass-1 tables (Q1)
  const u16* lc3;         // k_search only: the I4 (type 3) part [8][3][68], in the warp's shared memory during pass 2
};

ZW_HD u32 level_cost_at(const CostCtx& cc, int ctype, int n, int ctx, int v) {
  u32 fixed = ZW_TAB(kLevelFixedCosts)[imin(v, 2047)];
  u32 variable = 0;
  if (cc.level_cost) {
    int band = ZW_TAB(kEncBands)[n];
    variable = cc.level_cost[((ctype * 8 + band) * 3 + ctx) * 68 + imin(v, 67)];
  }
  return fixed + variable;
}

// Cost of coding `lv` (levels indexed 0..15 as given; the callers pass NATURAL order, Q7).
ZW_HD u32 residual_cost(const i32* lv, int ctype, int first, int ctx0, const CostCtx& cc) {
  int last = -1;
#pragma unroll
  for (int i = 0; i < 16; i++)
    if (lv[i] != 0) last = i;
  u32 p0 = cc.probs[((ctype * 8 + ZW_TAB(kEncBands)[first]) * 3 + ctx0) * 11];
  if (last < 0) return bit_cost(0, p0);
  u32 cost = ctx0 == 0 ? bit_cost(1, p0) : 0;
  int ctx = ctx0;
#pragma unroll
  for (int n = 0; n < 16; n++) {
    if (n >= first && n <= last) {
      int v = iabs(lv[n]);
      cost += level_cost_at(cc, ctype, n, ctx, v);
      if (n == last) {
        if (n < 15) {
          int nb = ZW_TAB(kEncBands)[n + 1];
          int nctx = v == 1 ? 1 : 2;
          cost += bit_cost(0, cc.probs[((ctype * 8 + nb) * 3 + nctx) * 11]);
        }
      }
      ctx = v >= 2 ? 2 : v;
    }
  }
  return cost;
}

// variable_level_cost (cost.rs:1425) + LevelCosts::calculate for one (type, band, ctx) row.
ZW_HD u32 variable_level_cost(int level, const u8* p) {
  if (level == 0) return 0;
  int idx = imin(level, 67) - 1;
  u32 pattern = ZW_TAB(kLevelCodes)[idx * 2 + 0];
  u32 bits = ZW_TAB(kLevelCodes)[idx * 2 + 1];
  u32 cost = 0;
  int i = 2;
  while (pattern != 0) {
    if (pattern & 1) cost += bit_cost((int)(bits & 1), p[i]);
    bits >>= 1;
    pattern >>= 1;
    i++;
  }
  return cost;
}
ZW_HD u16 level_cost_entry(const u8* p /*11 probs of (type,band,ctx)*/, int ctx, int v) {
  u32 cost0 = ctx > 0 ? bit_cost(1, p[0]) : 0;
  if (v == 0) return (u16)(bit_cost(0, p[1]) + cost0);
  u32 cost_base = bit_cost(1, p[1]) + cost0;
  return (u16)(cost_base + variable_level_cost(v, p));
}

// Trellis quantisation of one block.  coeffs: natural order DCT coefficients in, dequantised
// levels out; out: zig-zag levels.  Returns has_nz.  Scores are i64 like the reference.
ZW_HD bool trellis_quantize(i32* coeffs, i32* out, const Matrix& m, const u16* sharpen, u32 lambda, int first,
                            const CostCtx& cc, int ctype, int ctx0) {
  const i64 MAX_COST = (i64)0x3fffffffffffffffLL;  // i64::MAX / 2
  const i64 lam = (i64)lambda;
  i32 thresh = ((i32)m.q[1] * (i32)m.q[1]) / 4;
  int last = first - 1;
#pragma unroll
  for (int n = 0; n < 16; n++) {
    int j = ZW_TAB(kZigzag)[n];
    if (n >= first && coeffs[j] * coeffs[j] > thresh) last = n;
  }
  if (last < 15) last += 1;

  int best_n = -1, best_delta = 0, best_prev = 0;
  const u8* P = cc.probs + ctype * (8 * 3 * 11);
  const u16* LC = cc.level_cost + ctype * (8 * 3 * 68);
  int band0 = ZW_TAB(kEncBands)[first];
  i64 best_score = (i64)bit_cost(0, P[(band0 * 3 + ctx0) * 11]) * lam;  // skip: EOB at `first`
  i64 init_rate = ctx0 == 0 ? (i64)bit_cost(1, P[(band0 * 3 + ctx0) * 11]) : 0;
  // score state of the two nodes of the previous position + the cost row their level selects
  i64 sc0 = init_rate * lam, sc1 = sc0;
  const u16* row0 = LC + (band0 * 3 + ctx0) * 68;
  const u16* row1 = row0;
  u32 signs = 0, prev0 = 0, prev1 = 0;  // per-position bit masks
  i32 lvl0[16];

#pragma unroll
  for (int n = 0; n < 16; n++) {
    lvl0[n] = 0;
    if (n >= first && n <= last) {
      int j = ZW_TAB(kZigzag)[n];
      int k = j > 0;
      i32 q = m.q[k];
      u32 iq = m.iq[k];
      i32 c = coeffs[j];
      if (c < 0) signs |= 1u << n;
      i32 cs = iabs(c) + (i32)sharpen[j];
      i32 level0 = imin(quantdiv((u32)cs, iq, 0), 2047);                 // neutral bias: (0<<17 + 128)>>8 == 0
      i32 thresh_level = imin(quantdiv((u32)cs, iq, 1u << 16), 2047);    // bias 0x80: ((0x80<<17)+128)>>8 == 65536
      lvl0[n] = level0;
      i64 w = ZW_TAB(kWeightTrellis)[j];
      i64 orig_sq = (i64)(cs * cs);
      i64 ns0 = MAX_COST, ns1 = MAX_COST;
      const u16 *nrow0 = nullptr, *nrow1 = nullptr;
#pragma unroll
      for (int delta = 0; delta < 2; delta++) {
        i32 level = level0 + delta;
        int ctx = imin(level, 2);
        const u16* nrow = (n + 1 < 16) ? LC + (ZW_TAB(kEncBands)[n + 1] * 3 + ctx) * 68 : nullptr;
        if (delta == 0) nrow0 = nrow; else nrow1 = nrow;
        if (level > thresh_level) continue;  // level >= 0 always
        i32 ne = cs - level * q;
        i64 base = 256 * (w * ((i64)(ne * ne) - orig_sq));
        u32 fixed = ZW_TAB(kLevelFixedCosts)[level] + (level > 0 ? 256u : 0u);
        int lc = imin(level, 67);
        // predecessors without a cost row only exist past position 15, which is never a predecessor
        i64 s0 = sc0 + (i64)(fixed + row0[lc]) * lam;
        i64 s1 = sc1 + (i64)(fixed + row1[lc]) * lam;
        i64 cur;
        int bp;
        if (s1 < s0) { cur = s1 + base; bp = 1; } else { cur = s0 + base; bp = 0; }
        if (delta == 0) { ns0 = cur; if (bp) prev0 |= 1u << n; } else { ns1 = cur; if (bp) prev1 |= 1u << n; }
        if (level != 0 && cur < best_score) {
          i64 eob = n < 15 ? (i64)bit_cost(0, P[(ZW_TAB(kEncBands)[n + 1] * 3 + ctx) * 11]) : 0;
          i64 term = cur + eob * lam;
          if (term < best_score) { best_score = term; best_n = n; best_delta = delta; best_prev = bp; }
        }
      }
      sc0 = ns0; sc1 = ns1; row0 = nrow0; row1 = nrow1;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; i++)
    if (i >= first) { out[i] = 0; coeffs[i] = 0; }
  // note: for first == 1 the reference clears indices 1..15 in BOTH arrays (natural index for
  // coeffs, zig-zag index for out); index 0 coincides in both orders, so `i >= first` is exact.
  if (best_n < 0) return false;
  bool has_nz = false;
  int delta = best_delta;
#pragma unroll
  for (int n = 15; n >= 0; n--) {
    if (n <= best_n && n >= first) {
      int j = ZW_TAB(kZigzag)[n];
      i32 level = lvl0[n] + delta;
      if ((signs >> n) & 1) level = -level;
      out[n] = level;
      has_nz |= level != 0;
      coeffs[j] = level * (i32)m.q[j > 0];
      int pv = (n == best_n) ? best_prev : (int)(((delta ? prev1 : prev0) >> n) & 1);
      delta = pv;
    }
  }
  return has_nz;
}

// Token-statistics events of one coded block (record_coeffs): calls f(slot, bit) for every
// ProbaStats::record in reference order, slot = ((t*8+band)*3+ctx)*11+node.
template <class F>
ZW_HD void token_events(const i16* zz /*zig-zag levels*/, int t, int first, int ctx, F&& f) {
  int last = -1;
  for (int i = 0; i < 16; i++)
    if (zz[i] != 0) last = i;
  int eob = last + 1;
  if (eob <= first) {
    f(((t * 8 + ZW_TAB(kEncBands)[first]) * 3 + ctx) * 11 + 0, 0);
    return;
  }
  bool skip_eob = false;
  int n = first;
  while (n < eob) {
    int base = ((t * 8 + ZW_TAB(kEncBands)[n]) * 3 + ctx) * 11;
    int v = iabs((i32)zz[n]);
    n++;
    if (!skip_eob) f(base + 0, 1);
    if (v == 0) {
      f(base + 1, 0);
      skip_eob = true;
      ctx = 0;
      continue;
    }
    f(base + 1, 1);
    if (v == 1) {
      f(base + 2, 0);
      ctx = 1;
    } else {
      f(base + 2, 1);
      v = imin(v, 67);
      if (v <= 4) {
        f(base + 3, 0);
        if (v == 2) {
          f(base + 4, 0);
        } else {
          f(base + 4, 1);
          f(base + 5, v == 4);
        }
      } else if (v <= 10) {
        f(base + 3, 1);
        f(base + 6, 0);
        f(base + 7, v > 6);
      } else {
        f(base + 3, 1);
        f(base + 6, 1);
        if (v < 3 + (8 << 2)) {
          f(base + 8, 0);
          f(base + 9, v >= 3 + (8 << 1));
        } else {
          f(base + 8, 1);
          f(base + 10, v >= 3 + (8 << 3));
        }
      }
      ctx = 2;
    }
  }
  if (n < 16) f(((t * 8 + ZW_TAB(kEncBands)[n]) * 3 + ctx) * 11 + 0, 0);
}

}  // namespace zw
#endif  // ZW_COST_CUH
