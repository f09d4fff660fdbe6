// zw_search.cuh -- the per-macroblock mode search + final transform, both passes (sm_100a).
//
// Scheduling: persistent kernels.  One warp owns one macroblock ROW of one image at a time and
// walks it left to right; rows are handed out through a global ticket counter in an order in
// which row y-1 of an image always precedes row y, many images interleaved.  Row y may process
// macroblock x once row y-1 has finished x+1 (left / top / top-right dependency, SURVEY.md
// Appendix C), which it learns from a per-row progress counter (release / acquire pattern, see ld_flag).
// Left-neighbour state (borders, complexities, chroma diffusion errors) never leaves the warp.
//
// Pass 1 is split in two kernels because of a reference quirk: in pass 1 the
// chroma error-diffusion state `left_derr` is NOT reset at row starts, so chroma of row y depends
// on the END of row y-1 -- a raster-order chain.  Luma does not depend on chroma, so
//   k_search<1>  : luma only, wavefront-parallel (I16 / I4 search + final luma transform)
//   k_chroma1    : chroma search + transform + skip/complexity bookkeeping, one warp per image
//                  walking all macroblocks in raster order
//   k_search<2>  : luma + chroma, wavefront-parallel (pass 2 resets left_derr per row)
//
// Reference (file:line under /root/reference, src/encoder/vp8.rs unless noted):
//   choose_macroblock_info :2202   pick_best_intra16 :1504   pick_best_intra4 :1790
//   pick_best_uv :2050             transform_luma_block :2647  transform_luma_blocks_4x4 :2785
//   transform_chroma_blocks :3039  apply_chroma_error_diffusion :572  check_all_coeffs_zero :962
//   create_border_luma/chroma src/common/prediction.rs:15/:85, predictors :164-554
#ifndef ZW_SEARCH_CUH
#define ZW_SEARCH_CUH
#include "zw_types.cuh"

namespace zw {

constexpr int SEARCH_WARPS = 4;  // warps per CTA of the free-running wavefront kernels (each warp independent)
// Pass-2 luma kernel: the warps of a CTA step through the three phases of a macroblock (I16 search,
// I4 search, final transform) in lock step, one barrier per phase.  The kernel is bound by
// instruction supply (97 KB of code, every warp somewhere else in it: 42 % of the stall samples were
// "no instruction"); warps that share a scheduler now run the same code at the same time.  Measured:
// pass 2 78 -> 67 ms in spite of the idle time at the barriers.  0 = free-running like pass 1.
#ifndef ZW_LS_WARPS
#define ZW_LS_WARPS 5
#endif
constexpr int LS_WARPS = ZW_LS_WARPS;
#ifndef ZW_I4_LANES8
#define ZW_I4_LANES8 3  // I4 candidates with 8 lanes x 2 values per candidate (four per step) instead of 16 lanes x 1 (two per step):
                        // bit 0 = in pass 1, bit 1 = in pass 2.  Measured: pass 1 23.0 -> 20.6 ms; pass 2 needs the registers of
                        // 20 warps per SM (96 registers): 41.2 -> 39.3 ms with 10-warp CTAs, 38.8 ms with 5-warp CTAs (with 24 warps /
                        // 80 registers it loses: 42.3; CTA sizes that spread unevenly over the 4 schedulers lose too: 6 -> 42.8, 11 -> 43.0)
#endif
#ifndef ZW_LS_I4SYNC
#define ZW_LS_I4SYNC 0
#endif
#ifndef ZW_SEARCH_MIN_BLOCKS
#define ZW_SEARCH_MIN_BLOCKS 6   // CTAs per SM the wavefront kernels are register-budgeted for
#endif
#ifndef ZW_EXP_ROUNDS
#define ZW_EXP_ROUNDS 2
#endif
constexpr unsigned FULL = 0xffffffffu;
constexpr i64 I64_MAX = 0x7fffffffffffffffLL;

struct __align__(16) WarpScratch {
  u8 src_y[256];
  u8 src_u[64];
  u8 src_v[64];
  u8 yws[17 * 32];   // bordered luma work buffer (prediction.rs LUMA_STRIDE = 32)
  u8 uvws[9 * 32];   // bordered chroma: U at columns 0..8, V at columns 16..24
  u8 left_y[20];     // [0] corner, [1..16] left column
  u8 left_u[12];
  u8 left_v[12];
  i32 dcbuf[32];
  i16 nat[16][16];   // natural-order luma levels (I4 search winners / I16 simple quantisation)
  i32 coef[16][16];  // natural-order DCT coefficients handed to / returned by the cooperative trellis
  u8 cand_px[10][16];   // I4 search: reconstructed pixels of each evaluated candidate (by rank)
  i16 cand_lv[10][16];  // I4 search: natural-order levels of each evaluated candidate (by rank)
  u8 cand_mode[12];     // I4 search: mode with rank r
  u8 bmodes[16];
  u8 dtab[48];       // I4: the 23 distinct 3-tap edge filters of the current sub-block, its DC ([23]), its TM pixels ([32 + n])
  u32 eob_pack[16];  // per row: I4 end-of-block cost after position n: ctx 1 in the low half, ctx 2 in the high half
  u32 p0c[4];        // per row: I4 bit_cost(0, p0) for ctx0 = 0, 1, 2 and bit_cost(1, p0) for ctx0 = 0
  u32 eob_pack0[16]; // the same two for the I16 AC blocks (type 0, first coefficient 1: p0 of band 1)
  u32 p0c0[4];
  u8 nzflag[32];     // per-block non-zero flags (scratch)
  MbRecord rec;      // staged record
};

struct SearchShared {
  int4 lk[7][32];       // per-lane constants: butterflies of coop_fdct / coop_idct, coefficient bands (fill_lane_consts)
  int4 lk8[6][32];      // per-lane constants of the 8-lanes-per-candidate I4 evaluation (fill_lane_consts8)
  u8 pred_idx[10][16];  // (mode, pixel) -> index into WarpScratch::dtab
  u16 dtaps[32];        // lane k -> the three edge taps of dtab[k]
  WarpScratch w[SEARCH_WARPS];
};

// Progress flags (row y-1 -> row y of the same image), a release / acquire pattern of the PTX memory model:
//   producer: every lane stores its part of MbBottom / nz_after / derr2 / uvflags; __syncwarp() (bar.warp.sync orders
//             the participating lanes' accesses); then LANE 0 ALONE issues fence.release.gpu + st.relaxed.gpu of the
//             flag, so the fence is cumulative over what the other lanes stored before the barrier.
//   consumer: lane 0 polls with ld.relaxed.gpu; once the poll succeeds it issues ONE fence.acquire.gpu (not one per
//             poll), then __syncwarp() / a shuffle releases the other lanes, whose loads are ordered after it.
// Cost on sm_100a (cuobjdump -sass): fence.release.gpu = MEMBAR.ALL.GPU, fence.acquire.gpu = CCTL.IVALL -- an
// invalidation of the SM's whole L1, which also drops the rate tables other warps keep there.  -DZW_ACQUIRE_FENCE=0
// builds the variant without it (round 1's protocol).  That variant relies on three hardware properties instead of
// the memory model: (1) everything a waiting row reads from its producer is read with ld.global.cg, which never
// allocates in or hits L1, so no stale L1 line can be observed; (2) the producer's MEMBAR.ALL.GPU makes its data
// visible in L2 before the flag store is performed; (3) an SM issues a warp's instructions in order and the dependent
// loads sit behind a branch on the polled value.  tests/test_gpu_soak.py runs both builds against one hash set.
#ifndef ZW_ACQUIRE_FENCE
#define ZW_ACQUIRE_FENCE 1
#endif
__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_release() { asm volatile("fence.release.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acquire() {
#if ZW_ACQUIRE_FENCE
  asm volatile("fence.acquire.gpu;" ::: "memory");
#endif
}
__device__ __forceinline__ void st_flag(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// lane 0 of the producer, after the __syncwarp() that follows the warp's data stores
__device__ __forceinline__ void publish_flag(int* p, int v) {
  fence_release();
  st_flag(p, v);
}
__device__ __forceinline__ i64 shfl64(i64 v, int src) {
  int lo = __shfl_sync(FULL, (int)(v & 0xffffffff), src);
  int hi = __shfl_sync(FULL, (int)(v >> 32), src);
  return (i64)(((u64)(u32)hi << 32) | (u64)(u32)lo);
}
// Sums over aligned groups of 16 / 8 lanes: one redux.sync per group mask (integer adds: exact).
__device__ __forceinline__ unsigned lane_id() {
  unsigned l;
  asm("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}
__device__ __forceinline__ i64 shfl_up64_full(i64 v, int d) {
  int lo = __shfl_up_sync(FULL, (int)(v & 0xffffffff), d);
  int hi = __shfl_up_sync(FULL, (int)(v >> 32), d);
  return (i64)(((u64)(u32)hi << 32) | (u64)(u32)lo);
}
#ifdef ZW_USE_REDUX
__device__ __forceinline__ int red16_add(int v) {
  return __reduce_add_sync(0xffffu << (lane_id() & 16), v);
}
__device__ __forceinline__ int red8_add(int v) {
  return __reduce_add_sync(0xffu << (lane_id() & 24), v);
}
#else  // measured on B200: redux.sync on sub-warp masks is slower than the xor-shuffle butterfly
__device__ __forceinline__ int red16_add(int v) {
  v += __shfl_xor_sync(FULL, v, 8);
  v += __shfl_xor_sync(FULL, v, 4);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 1);
  return v;
}
__device__ __forceinline__ int red8_add(int v) {
  v += __shfl_xor_sync(FULL, v, 4);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 1);
  return v;
}
#endif

// 16x16 / 8x8 whole-block predictors evaluated per pixel from the bordered work buffer
// (predict_vpred/hpred/dcpred/tmpred, prediction.rs:164-324).  mode: 0 DC 1 V 2 H 3 TM.
__device__ __forceinline__ int pred_big(const u8* ws, int cofs, int mode, int x, int y, int dcv) {
  if (mode == 0) return dcv;
  if (mode == 1) return ws[cofs + 1 + x];
  if (mode == 2) return ws[(1 + y) * 32 + cofs];
  return clip255((int)ws[(1 + y) * 32 + cofs] + (int)ws[cofs + 1 + x] - (int)ws[cofs]);
}

// The same for a whole 4x4 block (bx,by) of the 16x16 / 8x8 prediction: 4 top + 4 left + corner
// loads, then selects.
__device__ __forceinline__ void pred_block(const u8* ws, int cofs, int mode, int bx, int by, int dcv, i32* pr) {
  // one formula for the four modes: clip255(T'[x] + L'[y] + base) with T' = T for V / TM (else 0),
  // L' = L for H / TM (else 0), base = dc for DC, -P for TM (else 0); the clip only acts for TM
  const bool useT = (mode & 1) != 0, useL = mode >= 2;
  i32 T[4], L[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    T[k] = useT ? (i32)ws[cofs + 1 + bx * 4 + k] : 0;
    L[k] = useL ? (i32)ws[(1 + by * 4 + k) * 32 + cofs] : 0;
  }
  const i32 base = mode == 0 ? dcv : (mode == 3 ? -(i32)ws[cofs] : 0);
#pragma unroll
  for (int k = 0; k < 16; k++) pr[k] = clip255(T[k & 3] + L[k >> 2] + base);
}

// 4x4 predictors (prediction.rs:326-554) as lookups.  Every pixel of the eight directional modes
// is one of 23 distinct 3-tap filters (x + 2y + z + 2) >> 2 of the 13 edge pixels
// e[0..12] = L3 L2 L1 L0 P A0..A7 (avg2(x,y) == taps (x,y,x); a copy == taps (x,x,x)): lane k
// evaluates filter k once per sub-block into WarpScratch::dtab, [23] holds the DC value, and a
// predicted pixel is dtab[d_pred_idx[mode][pixel]] (TM is computed per pixel).
// Same values as predict4_pixel() / ZW_PRED_TABLE_INIT in zw_prims.cuh (checked by tests/hostcheck).
__device__ const u16 d_dtaps[32] = ZW_DTAPS_INIT;
__device__ const u8 d_pred_idx[10][16] = ZW_PRED_IDX_INIT;

// Lane-private residual cost (get_residual_cost, cost.rs:1670-1729) in flat form.  The "context
// chain" is no chain: the context of position n is min(|level[n-1]|, 2), known up front, so the 16
// terms are independent loads (the rolled per-coefficient loop this replaces kept the levels in local
// memory and serialised 16 dependent iterations -- the latency of the pass-1 chroma chain).
// TYPE / FIRST are compile-time: the band of every position folds to a constant.
__device__ __forceinline__ constexpr int enc_band(int n) { return n < 4 ? n : (n == 4 ? 6 : (n == 5 ? 4 : (n == 6 ? 5 : (n < 15 ? 6 : (n == 15 ? 7 : 0))))); }
template <int TYPE, int FIRST>
__device__ __forceinline__ u32 residual_cost_flat(const i32* lv, int ctx0, const CostCtx& cc) {
  int last = -1, vlast = 0;
#pragma unroll
  for (int i = FIRST; i < 16; i++)
    if (lv[i] != 0) { last = i; vlast = iabs(lv[i]); }
  const u8* pp = cc.probs + TYPE * 264;
  const u32 p0 = pp[(enc_band(FIRST) * 3 + ctx0) * 11];
  const u16* lc = cc.level_cost ? cc.level_cost + TYPE * 1632 : nullptr;
  u32 cost = ctx0 == 0 ? bit_cost(1, p0) : 0;
#pragma unroll
  for (int n = FIRST; n < 16; n++) {
    const int v = iabs(lv[n]);
    const int ctx = n == FIRST ? ctx0 : imin(iabs(lv[n - 1]), 2);
    u32 c = ZW_TAB(kLevelFixedCosts)[imin(v, 2047)];
    if (lc) c += lc[(enc_band(n) * 3 + ctx) * 68 + imin(v, 67)];
    cost += n <= last ? c : 0u;
  }
  if (last < 15 && last >= 0) cost += bit_cost(0, pp[(ZW_TAB(kEncBands)[last + 1] * 3 + (vlast == 1 ? 1 : 2)) * 11]);
  return last < 0 ? bit_cost(0, p0) : cost;
}

// Per sub-block set-up of the predictor lookups: gathers the 13 edge pixels (one per lane), fills
// W.dtab and returns this lane's DC / TM predictions (n = lane & 15 is the lane's pixel).
struct Pred4 {
  i32 dc, tm;
};
__device__ __forceinline__ Pred4 pred4_prepare(WarpScratch& W, int x0, int y0, u32 taps, int lane) {
  // lane k < 4: L(3-k) = yws[(y0+3-k)][x0-1]; k = 4: P; k = 5..12: A(k-5) on the row above
  const int k = imin(lane, 12);
  const int off = k < 4 ? (y0 + 3 - k) * 32 + x0 - 1 : (y0 - 1) * 32 + x0 - 5 + k;
  const i32 e = W.yws[off];
  const i32 a = __shfl_sync(FULL, e, taps & 15), b = __shfl_sync(FULL, e, (taps >> 4) & 15), c = __shfl_sync(FULL, e, (taps >> 8) & 15);
  const bool in_dc = lane < 4 || (lane >= 5 && lane < 9);
  Pred4 r;
  r.dc = (__reduce_add_sync(FULL, in_dc ? e : 0) + 4) >> 3;
  W.dtab[lane] = (u8)(lane == 23 ? r.dc : ((a + 2 * b + c + 2) >> 2));
  const int n = lane & 15;
  const i32 l = __shfl_sync(FULL, e, 3 - (n >> 2)), t = __shfl_sync(FULL, e, 5 + (n & 3)), p = __shfl_sync(FULL, e, 4);
  r.tm = clip255(l + t - p);
  if (lane < 16) W.dtab[32 + lane] = (u8)r.tm;
  __syncwarp();
  return r;
}
__device__ __forceinline__ i32 pred4_get(const WarpScratch& W, const u8 (*pidx)[16], int mode, int n, const Pred4& p) {
  (void)p;
  return W.dtab[pidx[mode][n]];  // DC at [23], TM at [32 + n], the directional modes among [0..22]
}

// ---------------------------------------------------------------------------------------------
// Cooperative 4x4 primitives: the 16 lanes of a half-warp hold one pixel / coefficient each
// (natural order, n = lane & 15 = 4*row + col) and exchange operands with xor-shuffles, so one
// warp works on two blocks per step.  Same arithmetic as fdct4x4 / idct4x4 / residual_cost in
// zw_prims.cuh / zw_cost.cuh (reference transform.rs:35-79,176-207; cost.rs:1670-1729).
// ---------------------------------------------------------------------------------------------
// Per-lane constants of the butterflies below (SearchShared::lk, filled by fill_lane_consts): the four
// cases per stage (which output a lane produces) become one multiply-add form, so a stage costs two
// or three instructions instead of computing all cases and selecting.
//   lk[0] = {8*sA, A, B, C}      fdct rows:    t = 8p + 8sA*v;  r = (t*A + p2*B + C) >> S
//   lk[1] = {S, sC, A', B'}      fdct columns: t = p + sC*r;    o = ((t*A' + p2*B' + C') >> S') + (E & (p2 != 0))
//   lk[2] = {C', S', E, src}     src: lane holding this lane's natural-order output (un-permute)
//   lk[3] = {KA, KB, sg, s1}     idct pass:    t = ((v*KA) >> 16) + sg*(p + ((p*KB) >> 16));  r = s1*t + s2*p
//   lk[4] = {s2, KA', KB', sg'}  (primed: the horizontal pass, indexed by x instead of y)
//   lk[5] = {s1', s2', zig-zag position of natural index n, 0}
//   lk[6] = {band(n), band(n+1), zigzag(n), trellis weight of zigzag(n)}   n = lane & 15
__device__ __forceinline__ void fill_lane_consts(int4 (*lk)[32], int t) {
  const int x = t & 3, y = (t >> 2) & 3;
  // (selects, not indexed local arrays: keeps the kernels' stack frames small)
  auto A4 = [](int i) { return i == 0 ? 1 : (i == 1 ? -1 : 2217); };
  auto B4 = [](int i) { return i < 2 ? 1 : (i == 2 ? 5352 : -5352); };
  const int rx = ((x & 1) << 1) | (x >> 1), ry = ((y & 1) << 1) | (y >> 1);
  lk[0][t] = make_int4(x < 2 ? 8 : -8, A4(x), B4(x), x < 2 ? 0 : (x == 2 ? 14500 : 7500));
  lk[1][t] = make_int4(x < 2 ? 0 : 12, y < 2 ? 1 : -1, A4(y), B4(y));
  lk[2][t] = make_int4(y < 2 ? 7 : (y == 2 ? 12000 : 51000), y < 2 ? 4 : 16, y == 2 ? 1 : 0, ry * 4 + rx);
  // idct case tables by index i (y for the vertical pass, x for the horizontal one):
  //   i=0: t = v + p      i=2: t = p - v      i=1: t = mulA(v) - mulB(p)      i=3: t = mulA(v) + mulB(p)
  //   then                i=0: r = t + p      i=3: r = p - t      i=1: r = p + t      i=2: r = t - p
  auto KA = [](int i) { return (i & 1) ? 35468 : (i == 0 ? 65536 : -65536); };
  auto KB = [](int i) { return (i & 1) ? 20091 : 0; };
  auto SG = [](int i) { return i == 1 ? -1 : 1; };
  auto S1 = [](int i) { return i == 3 ? -1 : 1; };
  auto S2 = [](int i) { return i == 2 ? -1 : 1; };
  lk[3][t] = make_int4(KA(y), KB(y), SG(y), S1(y));
  lk[4][t] = make_int4(S2(y), KA(x), KB(x), SG(x));
  int zinv = 0;  // zig-zag position of natural index n = t & 15
  for (int z = 0; z < 16; z++) if (ZW_TAB(kZigzag)[z] == (t & 15)) zinv = z;
  lk[5][t] = make_int4(S1(x), S2(x), zinv, 0);
  // position n = t & 15: band of n, band of n + 1, natural index of zig-zag position n, its trellis weight
  const int n = t & 15, zz = ZW_TAB(kZigzag)[n];
  lk[6][t] = make_int4(ZW_TAB(kEncBands)[n], ZW_TAB(kEncBands)[n + 1], zz, (int)ZW_TAB(kWeightTrellis)[zz]);
}

__device__ __forceinline__ i32 coop_fdct(i32 v, int lane, const int4 (*lk)[32]) {
  const int4 k0 = lk[0][lane], k1 = lk[1][lane], k2 = lk[2][lane];
  // rows: lanes x=0..3 end up holding outputs 0,2,1,3 of their row
  i32 p = __shfl_xor_sync(FULL, v, 3);
  i32 t = v * k0.x + (p << 3);
  i32 p2 = __shfl_xor_sync(FULL, t, 1);
  const i32 r = (t * k0.y + p2 * k0.z + k0.w) >> k1.x;
  // columns: lanes y=0..3 end up holding outputs 0,2,1,3 of their column
  p = __shfl_xor_sync(FULL, r, 12);
  t = r * k1.y + p;
  p2 = __shfl_xor_sync(FULL, t, 4);
  const i32 o = ((t * k1.z + p2 * k1.w + k2.x) >> k2.y) + ((p2 != 0) ? k2.z : 0);
  // lane (x,y) holds coefficient (R[x], R[y]), R = {0,2,1,3} (an involution): un-permute
  return __shfl_sync(FULL, o, (lane & 16) | k2.w);
}

__device__ __forceinline__ i32 coop_idct(i32 v, int lane, const int4 (*lk)[32]) {
  const int4 k3 = lk[3][lane], k4 = lk[4][lane], k5 = lk[5][lane];
  // vertical pass: rows 0/2 form a1,b1; rows 1/3 form c1,d1
  i32 p = __shfl_xor_sync(FULL, v, 8);
  i32 t = ((v * k3.x) >> 16) + k3.z * (p + ((p * k3.y) >> 16));
  p = __shfl_xor_sync(FULL, t, 12);
  const i32 r = k3.w * t + k4.x * p;
  // horizontal pass with the final rounding
  p = __shfl_xor_sync(FULL, r, 2);
  t = ((r * k4.y) >> 16) + k4.w * (p + ((p * k4.z) >> 16));
  p = __shfl_xor_sync(FULL, t, 3);
  const i32 o = k5.x * t + k5.y * p;
  return (o + 4) >> 3;
}

// ---------------------------------------------------------------------------------------------
// I4 candidates, EIGHT lanes per candidate and two values per lane: all four method-4 candidates of a
// sub-block in one pass instead of two.  lane = cand * 8 + j, j = 2 * r + t:
//   pixel domain:       row r;            t = 0: columns (0, 1)     t = 1: columns (3, 2)
//   coefficient domain: row R[r] = {0,2,1,3}[r];  t = 0: columns (0, 2)     t = 1: columns (1, 3)
// In this layout the first stage of every butterfly pass is an exchange with lane ^ 1 (columns) or
// lane ^ 6 / lane ^ 2 (rows) and half of the arithmetic is in-lane.  Per-lane constants (lk8):
//   [0] = {s8, F0, F1, FC}  [1] = {G0, G1, GC, FS}   fdct rows:    u = 8 * recv + s8 * own; first = (u0*F0 + u1*F1 + FC) >> FS; second likewise with G
//   [2] = {sv, VA, VB, VC}  [3] = {VS, VE, KA, KB}   fdct columns: t = recv + sv * own;     out = ((t*VA + recv2*VB + VC) >> VS) + (VE & (recv2 != 0))
//   [4] = {sg, s2, HA, HB}                          idct: x = ((w*KA) >> 16) + sg * mulB'(recv);  y = recv2 + s2 * x;  first = g(y0) + f(y1), second = f(y0) - g(y1)
//   [5] = {sh, pixel / coefficient indices, bands / previous-row source lane, 0}
// Same arithmetic as fdct4x4 / idct4x4 (zw_prims.cuh).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fill_lane_consts8(int4 (*lk8)[32], int lane) {
  const int j = lane & 7, r = j >> 1, t = j & 1;
  const int R = ((r & 1) << 1) | (r >> 1);
  lk8[0][lane] = t ? make_int4(-8, 5352, 2217, 14500) : make_int4(8, 1, 1, 0);
  lk8[1][lane] = t ? make_int4(2217, -5352, 7500, 12) : make_int4(1, -1, 0, 0);
  lk8[2][lane] = make_int4(r < 2 ? 1 : -1, r == 1 ? -1 : (r == 0 ? 1 : 2217), r < 2 ? 1 : (r == 2 ? 5352 : -5352), r < 2 ? 7 : (r == 2 ? 12000 : 51000));
  lk8[3][lane] = make_int4(r < 2 ? 4 : 16, r == 2 ? 1 : 0, r == 0 ? 65536 : (r == 1 ? -65536 : 35468), r < 2 ? 0 : 20091);
  lk8[4][lane] = make_int4(r == 2 ? -1 : 1, r < 2 ? 1 : -1, t ? 35468 : 65536, t ? 20091 : 0);
  const int npx0 = 4 * r + (t ? 3 : 0), npx1 = 4 * r + (t ? 2 : 1);
  const int nc0 = 4 * R + (t ? 1 : 0), nc1 = 4 * R + (t ? 3 : 2);
  // lane (within the candidate) holding the last coefficient of the previous coefficient row: type B of row-lane Rinv[R - 1]
  const int Rm = R > 0 ? R - 1 : 0, rprev = ((Rm & 1) << 1) | (Rm >> 1), psrc = 2 * rprev + 1;
  lk8[5][lane] = make_int4(t ? -1 : 1, npx0 | (npx1 << 8) | (nc0 << 16) | (nc1 << 24),
                           (int)ZW_TAB(kEncBands)[nc0] | ((int)ZW_TAB(kEncBands)[nc1] << 8) | (psrc << 16), 0);
}
__device__ __forceinline__ void fdct8(i32& v0, i32& v1, int lane, const int4 (*lk8)[32]) {
  const int4 k0 = lk8[0][lane], k1 = lk8[1][lane], k2 = lk8[2][lane], k3 = lk8[3][lane];
  const i32 r0 = __shfl_xor_sync(FULL, v0, 1), r1 = __shfl_xor_sync(FULL, v1, 1);
  const i32 u0 = (r0 << 3) + k0.x * v0, u1 = (r1 << 3) + k0.x * v1;
  const i32 h0 = (u0 * k0.y + u1 * k0.z + k0.w) >> k1.w, h1 = (u0 * k1.x + u1 * k1.y + k1.z) >> k1.w;
  const i32 p0 = __shfl_xor_sync(FULL, h0, 6), p1 = __shfl_xor_sync(FULL, h1, 6);
  const i32 t0 = p0 + k2.x * h0, t1 = p1 + k2.x * h1;
  const i32 q0 = __shfl_xor_sync(FULL, t0, 2), q1 = __shfl_xor_sync(FULL, t1, 2);
  v0 = ((t0 * k2.y + q0 * k2.z + k2.w) >> k3.x) + (q0 != 0 ? k3.y : 0);
  v1 = ((t1 * k2.y + q1 * k2.z + k2.w) >> k3.x) + (q1 != 0 ? k3.y : 0);
}
__device__ __forceinline__ void idct8(i32& w0, i32& w1, int lane, const int4 (*lk8)[32]) {
  const int4 k3 = lk8[3][lane], k4 = lk8[4][lane];
  const int sh = lk8[5][lane].x;
  const i32 p0 = __shfl_xor_sync(FULL, w0, 2), p1 = __shfl_xor_sync(FULL, w1, 2);
  const i32 x0 = ((w0 * k3.z) >> 16) + k4.x * (p0 + ((p0 * k3.w) >> 16));
  const i32 x1 = ((w1 * k3.z) >> 16) + k4.x * (p1 + ((p1 * k3.w) >> 16));
  const i32 q0 = __shfl_xor_sync(FULL, x0, 6), q1 = __shfl_xor_sync(FULL, x1, 6);
  const i32 y0 = q0 + k4.y * x0, y1 = q1 + k4.y * x1;
  // f(x) = (x * HA) >> 16, g(x) = x + ((x * HB) >> 16): identity for t = 0, the two IDCT multipliers for t = 1
  const i32 f0 = (y0 * k4.z) >> 16, f1 = (y1 * k4.z) >> 16;
  const i32 g0 = y0 + ((y0 * k4.w) >> 16), g1 = y1 + ((y1 * k4.w) >> 16);
  const i32 first = g0 + f1, second = f0 - g1;
  const i32 e0 = __shfl_xor_sync(FULL, first, 1), e1 = __shfl_xor_sync(FULL, second, 1);
  w0 = (e0 + sh * first + 4) >> 3;
  w1 = (e1 + sh * second + 4) >> 3;
}
__device__ __forceinline__ int red8_max(int v) {
  v = max(v, __shfl_xor_sync(FULL, v, 1));
  v = max(v, __shfl_xor_sync(FULL, v, 2));
  v = max(v, __shfl_xor_sync(FULL, v, 4));
  return v;
}

// Y2 transforms with one coefficient per lane (n = lane & 15 = block index in raster order).  Both are
// the same +-1 butterflies (wht4x4 / iwht4x4 in zw_prims.cuh; transform.rs:116-158, :82-114) with
// different rounding: the lane pattern is that of coop_fdct (outputs land in 0,2,1,3 order per pass).
template <bool FORWARD>
__device__ __forceinline__ i32 coop_wht_any(i32 v, int lane) {
  const int x = lane & 3, y = (lane >> 2) & 3;
  // FORWARD runs rows then columns, the inverse columns then rows; the arithmetic per pass is the same
  const int m1a = FORWARD ? 3 : 12, m1b = FORWARD ? 1 : 4, m2a = FORWARD ? 12 : 3, m2b = FORWARD ? 4 : 1;
  const int i1 = FORWARD ? x : y, i2 = FORWARD ? y : x;
  i32 p = __shfl_xor_sync(FULL, v, m1a);
  i32 t = i1 < 2 ? v + p : p - v;          // a b c d
  p = __shfl_xor_sync(FULL, t, m1b);
  i32 r = (i1 == 0 || i1 == 2) ? t + p : (i1 == 1 ? p - t : t - p);  // y0 y2 y1 y3
  p = __shfl_xor_sync(FULL, r, m2a);
  t = i2 < 2 ? r + p : p - r;
  p = __shfl_xor_sync(FULL, t, m2b);
  i32 o = (i2 == 0 || i2 == 2) ? t + p : (i2 == 1 ? p - t : t - p);
  o = FORWARD ? (o + (o > 0 ? 1 : 0)) / 2 : (o + 3) >> 3;
  // lane (x,y) holds output (R[x], R[y]), R = {0,2,1,3}: un-permute
  const int rx = ((x & 1) << 1) | (x >> 1), ry = ((y & 1) << 1) | (y >> 1);
  return __shfl_sync(FULL, o, (lane & 16) | (ry * 4 + rx));
}

__device__ __forceinline__ int half_sum(int v) {  // sum over the 16 lanes of each half-warp
  return red16_add(v);
}

// residual_cost with one level per lane (natural order n = lane & 15).  Uniform per half-warp.
__device__ __forceinline__ u32 coop_residual_cost(i32 lv, int ctype, int first, int ctx0, const CostCtx& cc, int lane, bool& has_nz,
                                                  const int4 (*lk)[32]) {
  const int n = lane & 15, h = lane >> 4;
  const int4 kb = lk[6][lane];  // band(n), band(n + 1)
  const int v = iabs(lv);
  const u32 nzm = (__ballot_sync(FULL, lv != 0) >> (16 * h)) & 0xffffu;
  const int last = nzm ? 31 - __clz(nzm) : -1;
  has_nz = nzm != 0;
  const u8* pr = cc.probs + ctype * 264;  // [band][ctx][11]
  const u32 p0 = pr[(first /* band(0) = 0, band(1) = 1 */ * 3 + ctx0) * 11];
  const int pv = __shfl_up_sync(FULL, v, 1, 16);
  const int ctx = n == first ? ctx0 : imin(pv, 2);
  u32 c = 0;
  if (n >= first && n <= last) {
    c = ZW_TAB(kLevelFixedCosts)[imin(v, 2047)];
    if (cc.level_cost) c += cc.level_cost[ctype * 1632 + (kb.x * 3 + ctx) * 68 + imin(v, 67)];
    if (n == last && n < 15) c += bit_cost(0, pr[(kb.y * 3 + (v == 1 ? 1 : 2)) * 11]);
    if (n == first && ctx0 == 0) c += bit_cost(1, p0);
  }
  const u32 sum = (u32)half_sum((int)c);
  return last < 0 ? bit_cost(0, p0) : sum;
}

// The same for the I4 search (ctype 3, first 0), with everything that only depends on the image's
// probabilities precomputed once per row (WarpScratch::eob_pack / p0c): c0 = bit_cost(0, p0),
// c1 = bit_cost(1, p0) if ctx0 == 0 else 0.
__device__ __forceinline__ u32 coop_cost_i4(i32 lv, int ctx0, const CostCtx& cc, u32 eob_pack, u32 c0, u32 c1, int band, int lane,
                                            bool& has_nz) {
  const int n = lane & 15, h = lane >> 4;
  const int v = iabs(lv);
  const u32 nzm = (__ballot_sync(FULL, lv != 0) >> (16 * h)) & 0xffffu;
  const int last = nzm ? 31 - __clz(nzm) : -1;
  has_nz = nzm != 0;
  const int pv = __shfl_up_sync(FULL, v, 1, 16);
  const int ctx = n == 0 ? ctx0 : imin(pv, 2);
  u32 c = ZW_TAB(kLevelFixedCosts)[imin(v, 2047)];
  if (cc.level_cost) c += cc.lc3[(band * 3 + ctx) * 68 + imin(v, 67)];
  if (n == last) c += v == 1 ? (eob_pack & 0xffffu) : (eob_pack >> 16);  // eob_pack is 0 for n = 15
  if (n == 0) c += c1;
  if (n > last) c = 0;
  const u32 sum = (u32)half_sum((int)c);
  return last < 0 ? c0 : sum;
}

// ---------------------------------------------------------------------------------------------
// Cooperative trellis: the 16 lanes of a half-warp own the 16 zig-zag positions of one block.
// Same result as the serial trellis_quantize (zw_cost.cuh; reference cost.rs:788-1006), computed
// as a prefix scan: the 2-node Viterbi step of position n is a 2x2 matrix M_n in the (min,+)
// semiring, M_n[p][d] = rate(prev node p -> node d)*lambda + distortion(d); the node scores of all
// positions are the prefix products M_first (x) ... (x) M_n applied to the initial score.  Integer
// (min,+) products are exact and associative, so every score equals the sequential one; back
// pointers, the terminal choice (first minimum in (n, delta) order, strictly below the skip score)
// and the unwinding then follow the reference's tie rules.  Both half-warps always execute the
// function (shuffles / ballots are warp-wide); `active` gates the stores.
// ---------------------------------------------------------------------------------------------
// "Invalid" is a large finite score: up to 16 of them may add up inside the scan (2^58 * 16 = 2^62
// still fits i64) while every real score stays far below 2^50, so plain adds need no saturation.
constexpr i64 T_INF = (i64)1 << 58;
__device__ __forceinline__ i64 tmin(i64 a, i64 b) { return a < b ? a : b; }
__device__ __forceinline__ i64 shfl_up64_16(i64 v, int d) {
  int lo = __shfl_up_sync(FULL, (int)(v & 0xffffffff), d, 16);
  int hi = __shfl_up_sync(FULL, (int)(v >> 32), d, 16);
  return (i64)(((u64)(u32)hi << 32) | (u64)(u32)lo);
}
__device__ __forceinline__ i64 shfl_xor64(i64 v, int m) {
  int lo = __shfl_xor_sync(FULL, (int)(v & 0xffffffff), m);
  int hi = __shfl_xor_sync(FULL, (int)(v >> 32), m);
  return (i64)(((u64)(u32)hi << 32) | (u64)(u32)lo);
}

// coef: the block's 16 natural-order coefficients in shared memory (overwritten with the
// dequantised levels); zz_out: 16 zig-zag levels.  Returns has_nz (uniform inside the half-warp).
// c: the block's coefficient at this lane's zig-zag position; dq: its dequantised level (0 below `first`);
// eobp / p0c: the per-row end-of-block / first-branch cost packs of the block's type (see k_search).
__device__ __noinline__ bool trellis_half(bool active, i32 c, i32& dq, i16* zz_out, const Matrix& m, const u16* sharpen, u32 lambda, int first,
                             const CostCtx& cc, int ctype, int ctx0, int lane, const int4 (*lk)[32], const u32* eobp, const u32* p0c) {
  const int n = lane & 15, h = lane >> 4;
  const int4 kb = lk[6][lane];  // band(n), band(n + 1), zigzag(n), trellis weight
  const int j = kb.z;
  const int kq = j > 0;
  const i32 q = m.q[kq];
  const u32 iq = m.iq[kq];
  const i64 lam = (i64)lambda;
  const i32 thresh = ((i32)m.q[1] * (i32)m.q[1]) / 4;
  const u32 big = (__ballot_sync(FULL, n >= first && c * c > thresh) >> (16 * h)) & 0xffffu;
  int last = big ? 31 - __clz(big) : first - 1;
  if (last < 15) last += 1;
  const bool inrange = n >= first && n <= last;
  const bool sign = c < 0;
  const i32 cs = iabs(c) + (i32)sharpen[j];
  const i32 level0 = imin(quantdiv((u32)cs, iq, 0), 2047);
  const i32 thresh_level = imin(quantdiv((u32)cs, iq, 1u << 16), 2047);
  // Fast path: a terminal node needs level != 0.  If no position of either block of this warp can
  // reach level 1 the result is "all zero, no coefficients" without running the search.
  if (!__any_sync(FULL, inrange && thresh_level >= 1)) {
    if (active && n >= first) zz_out[n] = 0;
    dq = 0;
    return false;
  }
  const u16* LC = ctype == 3 ? cc.lc3 : cc.level_cost + ctype * (8 * 3 * 68);
  const int band = kb.x;
  i64 base[2];
  u32 fx[2];
  int lc[2], cx[2];
  bool valid[2];
  const i64 wgt = kb.w;
  const i64 orig_sq = (i64)(cs * cs);
#pragma unroll
  for (int d = 0; d < 2; d++) {
    const i32 level = level0 + d;
    valid[d] = level <= thresh_level;
    const i32 ne = cs - level * q;
    base[d] = 256 * (wgt * ((i64)(ne * ne) - orig_sq));
    fx[d] = (u32)ZW_TAB(kLevelFixedCosts)[imin(level, 2047)] + (level > 0 ? 256u : 0u);
    lc[d] = imin(level, 67);
    cx[d] = imin(level, 2);
  }
  // contexts of the previous position's two nodes (the initial context at `first`)
  const int pcx = __shfl_up_sync(FULL, cx[0] | (cx[1] << 2), 1, 16);
  int pcx0 = pcx & 3, pcx1 = pcx >> 2;
  if (n == first) { pcx0 = ctx0; pcx1 = ctx0; }
  if (n < first) { pcx0 = 0; pcx1 = 0; }
  i64 rate[2][2];  // [prev node][this node], already multiplied by lambda
#pragma unroll
  for (int d = 0; d < 2; d++) {
    rate[0][d] = (i64)(fx[d] + (u32)LC[(band * 3 + pcx0) * 68 + lc[d]]) * lam;
    rate[1][d] = (i64)(fx[d] + (u32)LC[(band * 3 + pcx1) * 68 + lc[d]]) * lam;
  }
  i64 P00, P01, P10, P11;  // prefix product, [virtual start node][this node]
  if (inrange) {
    P00 = valid[0] ? rate[0][0] + base[0] : T_INF;
    P01 = valid[1] ? rate[0][1] + base[1] : T_INF;
    P10 = valid[0] ? rate[1][0] + base[0] : T_INF;
    P11 = valid[1] ? rate[1][1] + base[1] : T_INF;
  } else {
    P00 = 0; P11 = 0; P01 = T_INF; P10 = T_INF;  // identity
  }
  // Hillis-Steele inclusive scan; positions beyond `last` are never read, so log2(last+1) steps do
  // (rolled loop, bound uniform across the warp: keeps the code small and skips dead steps)
  const int span = max(last, __shfl_xor_sync(FULL, last, 16));
#pragma unroll 1
  for (int d = 1; d <= span; d <<= 1) {
    const i64 L00 = shfl_up64_16(P00, d), L01 = shfl_up64_16(P01, d), L10 = shfl_up64_16(P10, d), L11 = shfl_up64_16(P11, d);
    if (n >= d) {
      const i64 n00 = tmin(L00 + P00, L01 + P10);
      const i64 n01 = tmin(L00 + P01, L01 + P11);
      const i64 n10 = tmin(L10 + P00, L11 + P10);
      const i64 n11 = tmin(L10 + P01, L11 + P11);
      P00 = n00; P01 = n01; P10 = n10; P11 = n11;
    }
  }
  const i64 init = (ctx0 == 0 ? (i64)p0c[3] : 0) * lam;  // bit_cost(1, p0(ctx 0))
  const i64 skip_score = (i64)p0c[ctx0] * lam;            // bit_cost(0, p0(ctx0))
  const u32 my_eob = eobp[n];                             // end-of-block cost after position n: ctx 1 | ctx 2 << 16
  i64 s[2];
  s[0] = init + tmin(P00, P10);
  s[1] = init + tmin(P01, P11);
  // back pointers (ties keep predecessor 0, cost.rs:927)
  // sp1 + r1 < sp0 + r0  <=>  sp1 - sp0 < r0 - r1 (exact: every term is far inside i64): only the difference travels
  i64 dsp = shfl_up64_16(s[1] - s[0], 1);
  if (n == first) dsp = 0;
  const bool bp0 = dsp < rate[0][0] - rate[1][0];
  const bool bp1 = dsp < rate[0][1] - rate[1][1];
  u32 bpm0 = (__ballot_sync(FULL, inrange && bp0) >> (16 * h)) & 0xffffu;
  u32 bpm1 = (__ballot_sync(FULL, inrange && bp1) >> (16 * h)) & 0xffffu;
  // terminal candidates in (n, delta) order: key = score * 32 + (2n + delta) keeps that order on ties
  i64 best = I64_MAX;
#pragma unroll
  for (int d = 0; d < 2; d++) {
    if (inrange && valid[d] && (level0 + d) != 0) {
      const i64 eob = (i64)(cx[d] == 1 ? (my_eob & 0xffffu) : (my_eob >> 16));  // 0 for n = 15; cx is 1 or 2 here
      const i64 key = (s[d] + eob * lam) * 32 + (n * 2 + d);
      best = tmin(best, key);
    }
  }
  {  // minimum over the 16 lanes of each half-warp: per half, a warp-wide signed minimum of the high words (the other half
     // masked out), then an unsigned minimum of the low words of the lanes that hold it -- four REDUX instead of eight shuffles
    const i32 khi = (i32)(best >> 32);
    const u32 klo = (u32)best;
    const i32 mh0 = __reduce_min_sync(FULL, h == 0 ? khi : 0x7fffffff), mh1 = __reduce_min_sync(FULL, h == 1 ? khi : 0x7fffffff);
    const u32 ml0 = __reduce_min_sync(FULL, (h == 0 && khi == mh0) ? klo : 0xffffffffu);
    const u32 ml1 = __reduce_min_sync(FULL, (h == 1 && khi == mh1) ? klo : 0xffffffffu);
    best = (i64)(((u64)(u32)(h ? mh1 : mh0) << 32) | (u64)(h ? ml1 : ml0));
  }
  const int bidx = (int)(best & 31);
  const bool have = (best >> 5) < skip_score;  // arithmetic shift == floor division by 32
  const int best_n = have ? (bidx >> 1) : -1;
  // unwind (uniform per half-warp): delta mask of the chosen path
  u32 dm = 0;
  {
    int dl = bidx & 1;
#pragma unroll 1
    for (int k = best_n; k >= first; k--) {
      dm |= (u32)dl << k;
      dl = (int)(((dl ? bpm1 : bpm0) >> k) & 1);
    }
  }
  i32 level = 0;
  if (have && inrange && n <= best_n) {
    level = level0 + (int)((dm >> n) & 1);
    if (sign) level = -level;
  }
  if (active && n >= first) zz_out[n] = (i16)level;
  dq = level * q;
  const u32 nzm = (__ballot_sync(FULL, level != 0) >> (16 * h)) & 0xffffu;
  return nzm != 0;
}

// ---------------------------------------------------------------------------------------------
// Source + border staging
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_luma_mb(WarpScratch& W, const ChunkParams& P, const u8* yp, int pw, int mbw,
                                             int mbx, int mby, u32 up_mb0, int lane) {
  if (lane < 16) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(yp + (size_t)(mby * 16 + lane) * pw + mbx * 16));
    *reinterpret_cast<uint4*>(&W.src_y[lane * 16]) = v;
  }
  if (mby == 0) {
    W.yws[lane] = 127;  // corner + 16 above + top-right: all 127 on the first row
  } else {
    const MbBottom* bt = &P.bottom[up_mb0 + mbx];
    if (lane < 16) W.yws[1 + lane] = __ldcg(&bt->y[lane]);
    else if (lane < 20)  // top-right 4 pixels: next MB's bottom row, or the replicated last pixel
      W.yws[1 + lane] = (mbx == mbw - 1) ? __ldcg(&bt->y[15]) : __ldcg(&(bt + 1)->y[lane - 16]);
  }
  __syncwarp();
  if (lane < 16) W.yws[(1 + lane) * 32] = (mbx == 0) ? 129 : W.left_y[1 + lane];
  if (lane == 0) W.yws[0] = (mby == 0) ? 127 : (mbx == 0 ? 129 : W.left_y[0]);
  if (lane < 12) {  // top-right copies for sub-block rows 1..3 (prediction.rs:49-54)
    const int r = 4 * (1 + lane / 4), c = 17 + (lane & 3);
    W.yws[r * 32 + c] = W.yws[c];
  }
  __syncwarp();
}

__device__ __forceinline__ void load_chroma_mb(WarpScratch& W, const ChunkParams& P, const u8* up, const u8* vp,
                                               int cwid, int mbx, int mby, u32 up_mb0, int lane) {
  if (lane < 8) {
    *reinterpret_cast<uint2*>(&W.src_u[lane * 8]) = __ldg(reinterpret_cast<const uint2*>(up + (size_t)(mby * 8 + lane) * cwid + mbx * 8));
  } else if (lane < 16) {
    const int r = lane - 8;
    *reinterpret_cast<uint2*>(&W.src_v[r * 8]) = __ldg(reinterpret_cast<const uint2*>(vp + (size_t)(mby * 8 + r) * cwid + mbx * 8));
  }
  if (mby == 0) {
    W.uvws[lane] = 127;
  } else {
    const MbBottom* bt = &P.bottom[up_mb0 + mbx];
    if (lane < 8) W.uvws[1 + lane] = __ldcg(&bt->u[lane]);
    else if (lane < 16) W.uvws[17 + (lane - 8)] = __ldcg(&bt->v[lane - 8]);
  }
  __syncwarp();
  if (lane < 8) {
    W.uvws[(1 + lane) * 32] = (mbx == 0) ? 129 : W.left_u[1 + lane];
    W.uvws[(1 + lane) * 32 + 16] = (mbx == 0) ? 129 : W.left_v[1 + lane];
  }
  if (lane == 0) {
    W.uvws[0] = (mby == 0) ? 127 : (mbx == 0 ? 129 : W.left_u[0]);
    W.uvws[16] = (mby == 0) ? 127 : (mbx == 0 ? 129 : W.left_v[0]);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Luma: I16 search, I4 search, final transform.  Leaves the reconstruction in W.yws, the coded
// zig-zag levels in W.rec.levels[0..16] and the sub-block modes in W.bmodes.
// ---------------------------------------------------------------------------------------------
struct LumaOut {
  bool use_i4;
  int mode16;
  u32 ynz;         // has_coeffs bit per coded luma block
  int y2nz;        // Y2 has_coeffs (I16 only)
  bool simple_nz;  // any SIMPLE-quantised luma level non-zero (skip test, Q13)
};

// LS: the CTA's warps run the three phases in lock step (a CTA barrier between them); `work` is
// false for a warp that only keeps the barriers company this round (no row / dependency not ready).
template <bool LS, bool L8>
__device__ LumaOut luma_mb(WarpScratch& W, const SearchShared& SH, const SegParams& SP, const CostCtx& cc, int i4_modes,
                           bool i4_always, bool trellis, int mbx, int mby, u32 in_top_nz, u32 in_left_nz, int lane, bool work) {
  LumaOut R;
  const int hb = lane >> 4, blk = lane & 15, bx = blk & 3, by = blk >> 2;
  const u8(*pidx)[16] = SH.pred_idx;
  // ===== pick_best_intra16 (vp8.rs:1504-1681) =====
  int dc16 = 0;
  int best16_mode = 0;
  u64 i16_score = 0;
  if (work) {
  {
    int s = 0;
    if (lane < 16) s = (mby != 0 ? W.yws[1 + lane] : 0) + (mbx != 0 ? W.yws[(1 + lane) * 32] : 0);
    s = red16_add(s);
    s = __shfl_sync(FULL, s, 0);
    const int shf = 3 + (mbx != 0) + (mby != 0);
    dc16 = (mbx == 0 && mby == 0) ? 128 : (s + (1 << (shf - 1))) >> shf;  // predict_dcpred :183
  }
  bool is_flat;  // is_flat_source_16 (cost.rs:177)
  int tsrc;      // TTransform of this lane's source block (cost.rs:73)
  {
    const int v0 = W.src_y[0];
    bool same = true;
    i32 px[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      px[k] = W.src_y[(by * 4 + (k >> 2)) * 16 + bx * 4 + (k & 3)];
      same &= px[k] == v0;
    }
    is_flat = __all_sync(FULL, same);
    tsrc = t_transform16(px, ZW_TAB(kWeightY));
  }
  i64 best16_score = I64_MAX;
  u32 best16_cc = 0, best16_mc = 0, best16_d = 0;
  i32 best16_sd = 0;
#pragma unroll 1
  for (int round = 0; round < ZW_EXP_ROUNDS; round++) {
    const int mode = round * 2 + hb;  // 0 DC, 1 V, 2 H, 3 TM (MODES order, vp8.rs:1509)
    const bool avail = !((mode == 1 && mby == 0) || (mode == 2 && mbx == 0) || (mode == 3 && (mbx == 0 || mby == 0)));
    i32 c[16], pr[16];
    pred_block(W.yws, 0, mode, bx, by, dc16, pr);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int x = bx * 4 + (k & 3), y = by * 4 + (k >> 2);
      c[k] = (i32)W.src_y[y * 16 + x] - pr[k];
    }
    fdct4x4(c);
    // Y2: the DC of this lane's block is coefficient `blk` (raster order) of its mode's Y2 block
    const i32 y2q = quantize_coeff(coop_wht_any<true>(c[0], lane), SP.y2, blk);
    bool y2_any;
    const u32 cost_y2 = coop_residual_cost(y2q, 1, 0, 0, cc, lane, y2_any, SH.lk);
    const i32 mydc = coop_wht_any<false>(dequantize(y2q, SP.y2, blk), lane);
    i32 lv[16];
    lv[0] = 0;
    int nzc = 0;
#pragma unroll
    for (int k = 1; k < 16; k++) { lv[k] = quantize_coeff(c[k], SP.y1, k); nzc += lv[k] != 0; }
    int cost_ac = (int)residual_cost_flat<0, 1>(lv, 0, cc);
#pragma unroll
    for (int k = 1; k < 16; k++) c[k] = dequantize(lv[k], SP.y1, k);
    c[0] = mydc;
    idct4x4(c);
    int sse = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int x = bx * 4 + (k & 3), y = by * 4 + (k >> 2);
      c[k] = clip255(pr[k] + c[k]);
      const int df = (i32)W.src_y[y * 16 + x] - c[k];
      sse += df * df;
    }
    int td = iabs(t_transform16(c, ZW_TAB(kWeightY)) - tsrc) >> 5;  // tdisto_4x4 (cost.rs:122)
    cost_ac = red16_add(cost_ac);
    sse = red16_add(sse);
    td = red16_add(td);
    nzc = red16_add(nzc);
    const u32 coeff_cost = cost_y2 + (u32)cost_ac;
    i32 sd = SP.tlambda > 0 ? (i32)(((i32)SP.tlambda * td + 128) >> 8) : 0;
    u32 dfin = (u32)sse;
    if (is_flat && nzc == 0) { dfin = dfin * 2; sd = sd * 2; }  // is_flat_coeffs(.., 16, 0)
    const u32 mode_cost = ZW_TAB(kFixedCostsI16)[mode];
    i64 score = ((i64)mode_cost + (i64)coeff_cost) * (i64)SP.lambda_i16 + 256 * ((i64)dfin + (i64)sd);
    if (!avail) score = I64_MAX;
#pragma unroll
    for (int h = 0; h < 2; h++) {  // the two halves in MODES order; strict < keeps the first
      const i64 s = shfl64(score, h * 16);
      const u32 ccst = __shfl_sync(FULL, coeff_cost, h * 16);
      const u32 dd = __shfl_sync(FULL, dfin, h * 16);
      const i32 ss = __shfl_sync(FULL, sd, h * 16);
      if (s < best16_score) {
        best16_score = s; best16_mode = round * 2 + h; best16_cc = ccst;
        best16_mc = ZW_TAB(kFixedCostsI16)[round * 2 + h]; best16_d = dd; best16_sd = ss;
      }
    }
  }
  {
    const i64 fs = ((i64)best16_mc + (i64)best16_cc) * (i64)SP.lambda_mode + 256 * ((i64)best16_d + (i64)best16_sd);
    i16_score = (u64)(fs > 0 ? fs : 0);
  }
  }  // work (I16 search)

  // ===== pick_best_intra4 (vp8.rs:1790-2036), gated as in choose_macroblock_info (:2210-2231) =====
  // Cooperative: each half-warp evaluates one prediction mode at a time, one pixel per lane.
  bool use_i4 = false;
  if (LS) __syncthreads();
  bool i4_go = work && i4_modes > 0 && (i4_always || i16_score > 211ull * (u64)SP.lambda_mode || best16_mode != 0);
  use_i4 = i4_go;
  {
    const int max_modes = i4_modes;
    const int n16 = lane & 15;
    const u32 taps = SH.dtaps[lane];
    // table indices of this lane's pixel for the modes 2r + hb of SSE steps r = 1..4, one per byte
    // prediction-SSE ranking on packed bytes: this lane's (mode slot, pixel row) = (lane >> 2, lane & 3); the four table
    // indices of that row of predicted pixels are one aligned word of pred_idx.  Step A: modes 0..7, step B: modes 8, 9.
    const u32 pkA = *reinterpret_cast<const u32*>(&pidx[lane >> 2][(lane & 3) * 4]);
    const u32 pkB = *reinterpret_cast<const u32*>(&pidx[8 + ((lane >> 2) & 1)][(lane & 3) * 4]);
    // this lane's quantiser entry, hoisted out of the sub-block loop (the compiler cannot: SP lives in global memory)
    const u32 lq_iq = SP.y1.iq[n16 > 0], lq_bias = SP.y1.bias[n16 > 0];
    const i32 lq_q = SP.y1.q[n16 > 0];
    const u32 lam_i4 = SP.lambda_i4, lam_mode = SP.lambda_mode;
    const u32 eobp = W.eob_pack[n16];
    const int band_n = SH.lk[6][lane].x;
    // (only used by the 8-lanes-per-candidate evaluation, L8)
    const int4 c8k = SH.lk8[5][lane];
    const int c8_cand = lane >> 3, c8_t = lane & 1, c8_r = (lane >> 1) & 3;
    const int c8_px0 = c8k.y & 255, c8_px1 = (c8k.y >> 8) & 255, c8_nc0 = (c8k.y >> 16) & 255, c8_nc1 = (c8k.y >> 24) & 255;
    const int c8_band0 = c8k.z & 255, c8_band1 = (c8k.z >> 8) & 255, c8_psrc = (c8k.z >> 16) & 255;
    const bool c8_dc = c8_nc0 == 0;
    const u32 c8_iq = SP.y1.iq[1], c8_bias = SP.y1.bias[1];
    const i32 c8_q = SP.y1.q[1];
    const u32 c8_iq0 = c8_dc ? SP.y1.iq[0] : c8_iq, c8_bias0 = c8_dc ? SP.y1.bias[0] : c8_bias;  // value 0 is the DC in one lane per group
    const i32 c8_q0 = c8_dc ? (i32)SP.y1.q[0] : c8_q;
    const u32 c8_eob0 = W.eob_pack[c8_nc0], c8_eob1 = W.eob_pack[c8_nc1];
    (void)eobp; (void)band_n; (void)lq_iq; (void)lq_bias; (void)lq_q; (void)c8_cand; (void)c8_t; (void)c8_r; (void)c8_px0; (void)c8_px1;
    (void)c8_band0; (void)c8_band1; (void)c8_psrc; (void)c8_iq0; (void)c8_bias0; (void)c8_q0; (void)c8_eob0; (void)c8_eob1;
    u64 running = 211ull * (u64)lam_mode;
    u32 total_mode_cost = 0;
    u32 tnz4 = 0, lnz4 = 0;  // MB-local non-zero context bits (Q7)
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
      if (LS && ZW_LS_I4SYNC) __syncthreads();  // optional finer lock step: one barrier per sub-block
      if (!i4_go) {
        if (LS && ZW_LS_I4SYNC) continue;
        break;
      }
      const int sbx = i & 3, sby = i >> 2, x0 = 1 + 4 * sbx, y0 = 1 + 4 * sby;
      const int top_ctx = sby == 0 ? 0 : W.bmodes[i - 4];
      const int left_ctx = sbx == 0 ? 0 : W.bmodes[i - 1];
      const int ctx0 = (sby == 0 ? 0 : (int)((tnz4 >> sbx) & 1)) + (sbx == 0 ? 0 : (int)((lnz4 >> sby) & 1));
      const i32 srcpx = W.src_y[(sby * 4 + (n16 >> 2)) * 16 + sbx * 4 + (n16 & 3)];
      const Pred4 P4 = pred4_prepare(W, x0, y0, taps, lane);
      // prediction SSE of the ten modes; the sort key (sse << 4 | m) of mode m: ascending key order == stable
      // ascending SSE order (sort_unstable_by_key is an insertion sort at this length, Q11)
      // Four lanes per mode, one row of four pixels each: the predicted row is four dtab bytes packed into a word,
      // |src - pred| per byte is one VABSDIFF4 and the row's sum of squares one DP4A; two xor-shuffles finish a mode.
      u32 skey;
      {
        const u32 src4 = *reinterpret_cast<const u32*>(&W.src_y[(sby * 4 + (lane & 3)) * 16 + sbx * 4]);
        const u32 pA = (u32)W.dtab[pkA & 255u] | ((u32)W.dtab[(pkA >> 8) & 255u] << 8) | ((u32)W.dtab[(pkA >> 16) & 255u] << 16) | ((u32)W.dtab[pkA >> 24] << 24);
        const u32 pB = (u32)W.dtab[pkB & 255u] | ((u32)W.dtab[(pkB >> 8) & 255u] << 8) | ((u32)W.dtab[(pkB >> 16) & 255u] << 16) | ((u32)W.dtab[pkB >> 24] << 24);
        const u32 dA = __vabsdiffu4(src4, pA), dB = __vabsdiffu4(src4, pB);
        u32 sA = __dp4a(dA, dA, 0u), sB = __dp4a(dB, dB, 0u);
        sA += __shfl_xor_sync(FULL, sA, 1); sB += __shfl_xor_sync(FULL, sB, 1);
        sA += __shfl_xor_sync(FULL, sA, 2); sB += __shfl_xor_sync(FULL, sB, 2);
        // key owners: lane 4m for mode m < 8; lanes 1 and 5 for modes 8 and 9 (every lane of a group holds the group's sum)
        skey = (lane & 3) == 0 ? ((sA << 4) | (u32)(lane >> 2)) : ((lane == 1 || lane == 5) ? ((sB << 4) | (u32)(8 + (lane >> 2))) : 0xffffffffu);
      }
#pragma unroll 1
      for (int r = 0; r < max_modes; r++) {  // extract the `max_modes` smallest keys in order
        const u32 kmin = __reduce_min_sync(FULL, skey);
        if (skey == kmin) { skey = 0xffffffffu; W.cand_mode[r] = (u8)(kmin & 15u); }
      }
      __syncwarp();
      u64 best_key = ~0ull;
      u32 best_sse = 0, best_rate = 0;
      int best_nz = 0;
      if constexpr (L8) {
      // evaluate the best `max_modes` candidates in rank order, FOUR per step (eight lanes and two values
      // per lane each, see fdct8); each lane group keeps the best of its own candidates (the rank in the
      // key makes keys unique: min == first best), merged once at the end
      {
        const int s0 = W.src_y[(sby * 4 + c8_r) * 16 + sbx * 4 + (c8_px0 & 3)], s1 = W.src_y[(sby * 4 + c8_r) * 16 + sbx * 4 + (c8_px1 & 3)];
        const u32 c0 = W.p0c[ctx0], c1 = ctx0 == 0 ? W.p0c[3] : 0u;
#pragma unroll 1
        for (int r = 0; r < max_modes; r += 4) {
          const int rank = r + c8_cand;
          const bool act = rank < max_modes;
          const int m = W.cand_mode[act ? rank : 0];
          const i32 pr0 = W.dtab[pidx[m][c8_px0]], pr1 = W.dtab[pidx[m][c8_px1]];
          i32 w0 = s0 - pr0, w1 = s1 - pr1;
          fdct8(w0, w1, lane, SH.lk8);
          // quantize_coeff: the DC entry only for coefficient 0 (value 0 of the first lane of a group)
          const i32 a0 = quantdiv((u32)iabs(w0), c8_iq0, c8_bias0);
          const i32 a1 = quantdiv((u32)iabs(w1), c8_iq, c8_bias);
          const i32 q0 = w0 < 0 ? -a0 : a0, q1 = w1 < 0 ? -a1 : a1;
          // get_residual_cost in natural order: position n's context is min(|level[n - 1]|, 2)
          // only min(|level|, 2) of the neighbouring positions is needed: both of a lane's values travel in one word
          const u32 m2 = (u32)imin(a0, 2) | ((u32)imin(a1, 2) << 2);
          const u32 em = __shfl_xor_sync(FULL, m2, 1);
          const u32 pm = __shfl_sync(FULL, m2, (lane & 24) | c8_psrc);  // last level of the previous coefficient row
          const int ctxa = c8_t ? (int)(em & 3u) : (c8_nc0 == 0 ? ctx0 : (int)(pm >> 2));
          const int ctxb = c8_t ? (int)(em >> 2) : (int)(em & 3u);
          const int last = red8_max(a1 != 0 ? c8_nc1 : (a0 != 0 ? c8_nc0 : -1));
          const bool nz = last >= 0;
          u32 t0 = ZW_TAB(kLevelFixedCosts)[imin(a0, 2047)], t1 = ZW_TAB(kLevelFixedCosts)[imin(a1, 2047)];
          if (cc.level_cost) {
            t0 += cc.lc3[(c8_band0 * 3 + ctxa) * 68 + imin(a0, 67)];
            t1 += cc.lc3[(c8_band1 * 3 + ctxb) * 68 + imin(a1, 67)];
          }
          if (c8_nc0 == last) t0 += a0 == 1 ? (c8_eob0 & 0xffffu) : (c8_eob0 >> 16);
          if (c8_nc1 == last) t1 += a1 == 1 ? (c8_eob1 & 0xffffu) : (c8_eob1 >> 16);
          if (c8_nc0 == 0) t0 += c1;
          u32 cs = (c8_nc0 <= last ? t0 : 0u) + (c8_nc1 <= last ? t1 : 0u);
          cs = (u32)red8_add((int)cs);
          const u32 coeff_cost = last < 0 ? c0 : cs;
          i32 d0 = q0 * c8_q0, d1 = q1 * c8_q;
          idct8(d0, d1, lane, SH.lk8);
          const i32 rec0 = clip255(pr0 + d0), rec1 = clip255(pr1 + d1);
          const i32 df0 = s0 - rec0, df1 = s1 - rec1;
          const u32 sse = (u32)red8_add(df0 * df0 + df1 * df1);
          if (act) {
            W.cand_px[rank][c8_px0] = (u8)rec0; W.cand_px[rank][c8_px1] = (u8)rec1;
            W.cand_lv[rank][c8_nc0] = (i16)q0;  W.cand_lv[rank][c8_nc1] = (i16)q1;
          }
          const u32 rate = ZW_TAB(kFixedCostsI4)[(top_ctx * 10 + left_ctx) * 10 + m] + coeff_cost;
          const u64 score = (u64)sse * 256ull + (u64)(rate & 0xffffu) * (u64)lam_i4;  // u16 truncation (Q8)
          const u64 key = act ? ((score << 4) | (u64)rank) : ~0ull;
          if (key < best_key) { best_key = key; best_sse = sse; best_rate = rate; best_nz = (int)nz; }
        }
        {  // merge the four lane groups (every lane of a group holds its group's best): the 64-bit minimum is two warp-wide
           // 32-bit minima (the high words, then the low words of the lanes holding the minimal high word); rank r was
           // evaluated by lane group r & 3, whose first lane hands over the winner's sse / rate / non-zero flag
          const u32 khi = (u32)(best_key >> 32), klo = (u32)best_key;
          const u32 mhi = __reduce_min_sync(FULL, khi);
          const u32 mlo = __reduce_min_sync(FULL, khi == mhi ? klo : 0xffffffffu);
          best_key = ((u64)mhi << 32) | (u64)mlo;
          const int wsrc = (int)(mlo & 3u) << 3;
          const u32 pack = __shfl_sync(FULL, (best_rate & 0xffffu) | ((u32)best_nz << 16), wsrc);  // only 16 bits of the rate are used (Q8)
          best_sse = __shfl_sync(FULL, best_sse, wsrc);
          best_rate = pack & 0xffffu;
          best_nz = (int)(pack >> 16);
        }
      }
      } else {
      // evaluate the best `max_modes` candidates in rank order, two per step; each half-warp keeps
      // the best of its own candidates (the rank in the key makes keys unique: min == first best)
#pragma unroll 1
      for (int r = 0; r < max_modes; r += 2) {
        const int rank = r + hb;
        const bool act = rank < max_modes;
        const int m = W.cand_mode[act ? rank : 0];
        const i32 pr = pred4_get(W, pidx, m, n16, P4);
        const i32 cf = coop_fdct(srcpx - pr, lane, SH.lk);
        const i32 ql = quantdiv((u32)iabs(cf), lq_iq, lq_bias);
        const i32 q = cf < 0 ? -ql : ql;  // quantize_coeff
        bool nz;
        const u32 coeff_cost = coop_cost_i4(q, ctx0, cc, eobp, W.p0c[ctx0], ctx0 == 0 ? W.p0c[3] : 0u, band_n, lane, nz);
        const i32 rec = clip255(pr + coop_idct(q * lq_q, lane, SH.lk));
        const i32 df = srcpx - rec;
        const u32 sse = (u32)half_sum(df * df);
        if (act) {
          W.cand_px[rank][n16] = (u8)rec;
          W.cand_lv[rank][n16] = (i16)q;
        }
        const u32 rate = ZW_TAB(kFixedCostsI4)[(top_ctx * 10 + left_ctx) * 10 + m] + coeff_cost;
        const u64 score = (u64)sse * 256ull + (u64)(rate & 0xffffu) * (u64)lam_i4;  // u16 truncation (Q8)
        const u64 key = act ? ((score << 4) | (u64)rank) : ~0ull;
        if (key < best_key) { best_key = key; best_sse = sse; best_rate = rate; best_nz = (int)nz; }
      }
      {  // merge the two halves
        const u64 k2 = (u64)shfl_xor64((i64)best_key, 16);
        const u32 s2 = __shfl_xor_sync(FULL, best_sse, 16), r2 = __shfl_xor_sync(FULL, best_rate, 16);
        const int z2 = __shfl_xor_sync(FULL, best_nz, 16);
        if (k2 < best_key) { best_key = k2; best_sse = s2; best_rate = r2; best_nz = z2; }
      }
      }
      __syncwarp();
      const int wrank = (int)(best_key & 15);
      const int wmode = W.cand_mode[wrank];
      if (lane < 16) {
        W.yws[(y0 + (lane >> 2)) * 32 + x0 + (lane & 3)] = W.cand_px[wrank][lane];
        W.nat[i][lane] = W.cand_lv[wrank][lane];
      }
      if (lane == 0) W.bmodes[i] = (u8)wmode;
      __syncwarp();
      tnz4 = (tnz4 & ~(1u << sbx)) | ((u32)best_nz << sbx);
      lnz4 = (lnz4 & ~(1u << sby)) | ((u32)best_nz << sby);
      total_mode_cost += ZW_TAB(kFixedCostsI4)[(top_ctx * 10 + left_ctx) * 10 + wmode];
      running += (u64)best_sse * 256ull + (u64)(best_rate & 0xffffu) * (u64)lam_mode;
      if (running >= i16_score || total_mode_cost > 16384u) { use_i4 = false; i4_go = false; }
    }
  }

  // ===== final luma transform -> coded levels + reconstruction =====
  if (LS) __syncthreads();
  bool any_simple_nz = false;
  u32 ynz = 0;
  int y2nz = 0;
  if (!work) {
  } else if (!use_i4) {
    // ---- transform_luma_block (vp8.rs:2647-2780) ----
    i32 c[16], pr[16];
    pred_block(W.yws, 0, best16_mode, bx, by, dc16, pr);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int x = bx * 4 + (k & 3), y = by * 4 + (k >> 2);
      c[k] = (i32)W.src_y[y * 16 + x] - pr[k];
    }
    fdct4x4(c);
    // both half-warps hold the same 16 blocks here, so the cooperative Y2 transforms agree in both
    const i32 y2q = quantize_coeff(coop_wht_any<true>(c[0], lane), SP.y2, blk);
    y2nz = (__ballot_sync(FULL, y2q != 0) & 0xffffu) != 0;
    if (lane < 16) {
      int zp = 0;  // zig-zag position of natural index `lane`
#pragma unroll
      for (int k = 0; k < 16; k++) if (ZW_TAB(kZigzag)[k] == lane) zp = k;
      W.rec.levels[0][zp] = (i16)y2q;
    }
    const i32 mydc = coop_wht_any<false>(dequantize(y2q, SP.y2, blk), lane);
    any_simple_nz |= y2nz != 0;
    bool my_simple = false;
    i32 lv[16];
    lv[0] = 0;
#pragma unroll
    for (int k = 1; k < 16; k++) { lv[k] = quantize_coeff(c[k], SP.y1, k); my_simple |= lv[k] != 0; }
    if (lane >= 16) my_simple = false;
    if (!trellis) {
      if (lane < 16) {
#pragma unroll
        for (int k = 1; k < 16; k++) c[k] = dequantize(lv[k], SP.y1, k);
#pragma unroll
        for (int k = 0; k < 16; k++) W.nat[lane][k] = (i16)lv[k];
        W.nzflag[lane] = my_simple;
      }
      __syncwarp();
      if (lane < 16) {
#pragma unroll
        for (int k = 0; k < 16; k++) W.rec.levels[1 + lane][k] = W.nat[lane][ZW_TAB(kZigzag)[k]];
      }
    } else {
      // trellis with the nz context chained in raster order (:2685-2728): blocks on one anti-diagonal
      // are independent, so sweep the diagonals two blocks (one per half-warp) at a time
      if (lane < 4) {
        W.nzflag[16 + lane] = (in_top_nz >> (1 + lane)) & 1;
        W.nzflag[20 + lane] = (in_left_nz >> (1 + lane)) & 1;
      }
      if (lane < 16) {
#pragma unroll
        for (int k = 0; k < 16; k++) W.coef[lane][k] = c[k];
      }
      __syncwarp();
      {
        // rounds: pairs of blocks from the same diagonal (-1 = idle half)
        const int r0[10] = {0, 1, 2, 8, 3, 9, 7, 13, 11, 15};
        const int r1[10] = {-1, 4, 5, -1, 6, 12, 10, -1, 14, -1};
#pragma unroll 1
        for (int r = 0; r < 10; r++) {
          const int b = hb ? r1[r] : r0[r];
          const bool act = b >= 0;
          const int bb = act ? b : 0;
          const int tbx = bb & 3, tby = bb >> 2;
          const int ctx0 = imin((int)W.nzflag[20 + tby] + (int)W.nzflag[16 + tbx], 2);
          const int zj = SH.lk[6][lane].z;  // natural index of this lane's zig-zag position
          i32 dqv;
          const bool nz = trellis_half(act, W.coef[bb][zj], dqv, W.rec.levels[1 + bb], SP.y1, SP.sharpen, SP.lambda_trellis_i16, 1, cc, 0, ctx0,
                                       lane, SH.lk, W.eob_pack0, W.p0c0);
          if (act && (lane & 15) >= 1) W.coef[bb][zj] = dqv;
          __syncwarp();
          if (act && (lane & 15) == 0) {
            W.nzflag[bb] = nz;
            W.nzflag[16 + tbx] = nz;
            W.nzflag[20 + tby] = nz;
          }
          __syncwarp();
        }
      }
      if (lane < 16) {
#pragma unroll
        for (int k = 1; k < 16; k++) c[k] = W.coef[lane][k];
      }
    }
    if (lane < 16) {
      c[0] = mydc;
      idct4x4(c);
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const int x = bx * 4 + (k & 3), y = by * 4 + (k >> 2);
        W.yws[(1 + y) * 32 + 1 + x] = (u8)clip255(pr[k] + c[k]);
      }
    }
    any_simple_nz |= __any_sync(FULL, my_simple);
    __syncwarp();
    ynz = __ballot_sync(FULL, lane < 16 && W.nzflag[lane] != 0) & 0xffffu;
  } else if (!trellis) {
    // ---- transform_luma_blocks_4x4 without trellis == what the search already produced ----
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 16; k++) W.rec.levels[0][k] = 0;
    }
    if (lane < 16) {
      bool nz = false;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const i16 v = W.nat[lane][ZW_TAB(kZigzag)[k]];
        W.rec.levels[1 + lane][k] = v;
        nz |= v != 0;
      }
      W.nzflag[lane] = nz;
    }
    __syncwarp();
    ynz = __ballot_sync(FULL, lane < 16 && W.nzflag[lane] != 0) & 0xffffu;
    any_simple_nz |= ynz != 0;
  } else {
    // ---- transform_luma_blocks_4x4 with trellis (vp8.rs:2785-2916).  The reference walks the 16
    //      sub-blocks in raster order; block (x, y) only needs its left, top and top-right neighbours
    //      (prediction, non-zero contexts), so blocks with equal x + 2y are independent: ten rounds, the
    //      two half-warps taking one block each in six of them (same results, 16 -> 10 steps) ----
    u32 tnz = (in_top_nz >> 1) & 15, lnz = (in_left_nz >> 1) & 15;
    bool simple_any = false;
    const int n16 = lane & 15;
    const u32 tq_iq = SP.y1.iq[n16 > 0], tq_bias = SP.y1.bias[n16 > 0];
    const int zz_nat = SH.lk[6][lane].z;  // natural index of zig-zag position n16
    const int zz_pos = SH.lk[5][lane].z;  // zig-zag position of natural index n16
    const int hsel = lane & 16;
#pragma unroll 1
    for (int rd = 0; rd < 10; rd++) {
      // round -> block of half 0 / half 1 (-1: idle): {0,-} {1,-} {2,4} {3,5} {6,8} {7,9} {10,12} {11,13} {14,-} {15,-}
      const int ba = rd < 4 ? rd : (rd < 8 ? 6 + ((rd - 4) >> 1) * 4 + (rd & 1) : 6 + rd);
      const int bb = (rd >= 2 && rd < 8) ? ba + 2 : -1;
      const int blk_i = hb ? bb : ba;
      const bool act = blk_i >= 0;
      const int i = act ? blk_i : ba;  // an idle half shadows half 0's block (its stores are gated)
      const int sbx = i & 3, sby = i >> 2, x0 = 1 + 4 * sbx, y0 = 1 + 4 * sby;
      const int bmode = W.bmodes[i];
      // the one predictor this block uses, straight from its 13 edge pixels (lane k of the half holds e[k])
      i32 pr;
      {
        const int k = imin(n16, 12);
        const int off = k < 4 ? (y0 + 3 - k) * 32 + x0 - 1 : (y0 - 1) * 32 + x0 - 5 + k;
        const i32 e = W.yws[off];
        const u32 tp = SH.dtaps[pidx[bmode][n16]];
        const i32 ta = __shfl_sync(FULL, e, hsel | (tp & 15)), tb = __shfl_sync(FULL, e, hsel | ((tp >> 4) & 15)),
                  tc = __shfl_sync(FULL, e, hsel | ((tp >> 8) & 15));
        pr = (ta + 2 * tb + tc + 2) >> 2;
        if (__any_sync(FULL, bmode < 2)) {  // DC / TM (uniform branch)
          const bool in_dc = n16 < 4 || (n16 >= 5 && n16 < 9);
          const i32 dc = (half_sum(in_dc ? e : 0) + 4) >> 3;
          const i32 l = __shfl_sync(FULL, e, hsel | (3 - (n16 >> 2))), t = __shfl_sync(FULL, e, hsel | (5 + (n16 & 3))),
                    pp = __shfl_sync(FULL, e, hsel | 4);
          if (bmode == 0) pr = dc;
          if (bmode == 1) pr = clip255(l + t - pp);
        }
      }
      const i32 cf = coop_fdct((i32)W.src_y[(sby * 4 + (n16 >> 2)) * 16 + sbx * 4 + (n16 & 3)] - pr, lane, SH.lk);
      if (act) simple_any |= quantdiv((u32)iabs(cf), tq_iq, tq_bias) != 0;  // quantize_coeff != 0
      const int ctx0 = imin((int)((lnz >> sby) & 1) + (int)((tnz >> sbx) & 1), 2);
      // coefficients travel in registers: natural order -> zig-zag order and back by shuffle
      const i32 czz = __shfl_sync(FULL, cf, hsel | zz_nat);
      i32 dqz;
      const bool nzh = trellis_half(act, czz, dqz, W.rec.levels[1 + i], SP.y1, SP.sharpen, SP.lambda_trellis_i4, 0, cc, 3, ctx0, lane, SH.lk,
                                    W.eob_pack, W.p0c);
      const i32 dqn = __shfl_sync(FULL, dqz, hsel | zz_pos);
      const i32 rec = clip255(pr + coop_idct(dqn, lane, SH.lk));
      if (act) W.yws[(y0 + (n16 >> 2)) * 32 + x0 + (n16 & 3)] = (u8)rec;
      // contexts: both halves learn both results (blocks of one round differ in x and in y)
      const u32 nza = (u32)(__shfl_sync(FULL, (int)nzh, 0) != 0), nzb = (u32)(__shfl_sync(FULL, (int)nzh, 16) != 0);
      {
        const int ax = ba & 3, ay = ba >> 2;
        tnz = (tnz & ~(1u << ax)) | (nza << ax);
        lnz = (lnz & ~(1u << ay)) | (nza << ay);
        ynz |= nza << ba;
        if (bb >= 0) {
          const int bx2 = bb & 3, by2 = bb >> 2;
          tnz = (tnz & ~(1u << bx2)) | (nzb << bx2);
          lnz = (lnz & ~(1u << by2)) | (nzb << by2);
          ynz |= nzb << bb;
        }
      }
      __syncwarp();
    }
    simple_any = __any_sync(FULL, simple_any);
    any_simple_nz |= simple_any;
  }
  __syncwarp();
  R.use_i4 = use_i4;
  R.mode16 = best16_mode;
  R.ynz = ynz;
  R.y2nz = y2nz;
  R.simple_nz = any_simple_nz;
  return R;
}

// ---------------------------------------------------------------------------------------------
// Chroma: pick_best_uv (vp8.rs:2050-2200) + transform_chroma_blocks with DC error diffusion
// (:3039-3121, :572-647).  Leaves the reconstruction in W.uvws and the coded zig-zag levels in
// W.rec.levels[17..24]; updates the packed diffusion state.
// ---------------------------------------------------------------------------------------------
struct ChromaOut {
  int uv_mode;
  u32 uvnz;  // has_coeffs bit per chroma block (== simple-quantised non-zero)
};

__device__ ChromaOut chroma_mb(WarpScratch& W, const SegParams& SP, const CostCtx& cc, int mbx, int mby, u32& left_derr,
                               u32& top_derr, int lane) {
  ChromaOut R;
  int dcU, dcV;
  {
    int su = 0, sv = 0;
    if (lane < 8) {
      su = (mby != 0 ? W.uvws[1 + lane] : 0) + (mbx != 0 ? W.uvws[(1 + lane) * 32] : 0);
      sv = (mby != 0 ? W.uvws[17 + lane] : 0) + (mbx != 0 ? W.uvws[(1 + lane) * 32 + 16] : 0);
    }
    su = red8_add(su); sv = red8_add(sv);
    su = __shfl_sync(FULL, su, 0); sv = __shfl_sync(FULL, sv, 0);
    const int shf = 2 + (mbx != 0) + (mby != 0);
    dcU = (mbx == 0 && mby == 0) ? 128 : (su + (1 << (shf - 1))) >> shf;
    dcV = (mbx == 0 && mby == 0) ? 128 : (sv + (1 << (shf - 1))) >> shf;
  }
  const int b8 = lane & 7, ch = b8 >> 2, cbx = b8 & 1, cby = (b8 >> 1) & 1, cofs = ch * 16;
  const u8* src = ch ? W.src_v : W.src_u;
  int uv_mode;
  // The final transform reuses what the winning mode's lanes computed in the search: prediction, raw DC
  // and the simple-quantised AC levels (transform_chroma_blocks quantises the same coefficients the same
  // way; only the DCs change through the error diffusion) -- fetched with shuffles instead of a second
  // prediction + FDCT + quantisation.
  i32 c[16], pr[16], q[16];
  {
    const int mode = lane >> 3;  // lane = mode*8 + block (0..3 U, 4..7 V); MODES order DC,V,H,TM
    const bool avail = !((mode == 1 && mby == 0) || (mode == 2 && mbx == 0) || (mode == 3 && (mbx == 0 || mby == 0)));
    pred_block(W.uvws, cofs, mode, cbx, cby, ch ? dcV : dcU, pr);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int x = cbx * 4 + (k & 3), y = cby * 4 + (k >> 2);
      c[k] = (i32)src[y * 8 + x] - pr[k];
    }
    fdct4x4(c);
    const i32 rawdc = c[0];
    int nzac = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) { q[k] = quantize_coeff(c[k], SP.uv, k); if (k > 0) nzac += q[k] != 0; }
    int cost = (int)residual_cost_flat<2, 0>(q, 0, cc);
#pragma unroll
    for (int k = 0; k < 16; k++) c[k] = dequantize(q[k], SP.uv, k);
    idct4x4(c);
    int sse = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int x = cbx * 4 + (k & 3), y = cby * 4 + (k >> 2);
      const int df = (i32)src[y * 8 + x] - clip255(pr[k] + c[k]);
      sse += df * df;
    }
    cost = red8_add(cost); sse = red8_add(sse); nzac = red8_add(nzac);
    const u32 penalty = (mode > 0 && nzac <= 2) ? 140u * 8u : 0u;  // is_flat_coeffs(.., 8, 2) -> FLATNESS_PENALTY*8
    i64 score = ((i64)ZW_TAB(kFixedCostsUV)[mode] + (i64)cost + (i64)penalty) * (i64)SP.lambda_uv + 256 * (i64)sse;
    if (!avail) score = I64_MAX;
    i64 best = I64_MAX;
    uv_mode = 0;
#pragma unroll
    for (int mm = 0; mm < 4; mm++) {
      const i64 s = shfl64(score, mm * 8);
      if (s < best) { best = s; uv_mode = mm; }
    }
    // ---- transform_chroma_blocks: lane b (< 8) takes over block b of the winning mode ----
    const int from = uv_mode * 8 + b8;
    c[0] = __shfl_sync(FULL, rawdc, from);
#pragma unroll
    for (int k = 1; k < 16; k++) q[k] = __shfl_sync(FULL, q[k], from);
#pragma unroll
    for (int k = 0; k < 16; k++) pr[k] = __shfl_sync(FULL, pr[k], from);
  }
  if (lane < 8) W.dcbuf[lane] = c[0];
  __syncwarp();
  // error diffusion of the 8 DCs, computed redundantly by every lane (tiny, strictly serial)
  i32 ndc[8];
  u32 new_left = 0, new_top = 0;
#pragma unroll
  for (int cch = 0; cch < 2; cch++) {
    const i32 qd = SP.uv.q[0];
    const u32 iq = SP.uv.iq[0], bias = SP.uv.bias[0];
    const i32 t0 = (i8)(top_derr >> (16 * cch)), t1 = (i8)(top_derr >> (16 * cch + 8));
    const i32 l0 = (i8)(left_derr >> (16 * cch)), l1 = (i8)(left_derr >> (16 * cch + 8));
    i32 errs[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const i32 te = b == 0 ? t0 : (b == 1 ? t1 : (b == 2 ? errs[0] : errs[1]));
      const i32 le = b == 0 ? l0 : (b == 1 ? errs[0] : (b == 2 ? l1 : errs[2]));
      i32 dc = W.dcbuf[cch * 4 + b];
      dc += (7 * te + 8 * le) >> 3;
      ndc[cch * 4 + b] = dc;
      const bool sign = dc < 0;
      const u32 a = (u32)iabs(dc);
      const i32 level = a > SP.uv_dc_zthresh ? (i32)((a * iq + bias) >> 17) : 0;
      const i32 err = (i32)a - level * qd;
      const i32 se = (sign ? -err : err) >> 1;
      errs[b] = imin(imax(se, -127), 127);
    }
    const i32 nl0 = errs[1], nl1 = (3 * errs[3]) >> 2;
    const i32 nt0 = errs[2], nt1 = errs[3] - nl1;
    new_left |= ((u32)(u8)(i8)nl0 | ((u32)(u8)(i8)nl1 << 8)) << (16 * cch);
    new_top |= ((u32)(u8)(i8)nt0 | ((u32)(u8)(i8)nt1 << 8)) << (16 * cch);
  }
  __syncwarp();
  left_derr = new_left;
  top_derr = new_top;
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (k == lane) c[0] = ndc[k];
    q[0] = quantize_coeff(c[0], SP.uv, 0);
    bool nz = false;
    constexpr int ZZ[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};  // kZigzag: levels go to the record straight from registers
#pragma unroll
    for (int k = 0; k < 16; k++) {
      nz |= q[k] != 0;
      c[k] = dequantize(q[k], SP.uv, k);
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
      reinterpret_cast<u32*>(W.rec.levels[17 + lane])[k] = (u32)(u16)(i16)q[ZZ[2 * k]] | ((u32)(u16)(i16)q[ZZ[2 * k + 1]] << 16);
    idct4x4(c);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int x = cbx * 4 + (k & 3), y = cby * 4 + (k >> 2);
      W.uvws[(1 + y) * 32 + cofs + 1 + x] = (u8)clip255(pr[k] + c[k]);
    }
    W.nzflag[lane] = nz;
  }
  __syncwarp();
  R.uvnz = __ballot_sync(FULL, lane < 8 && W.nzflag[lane] != 0) & 0xffu;
  R.uv_mode = uv_mode;
  __syncwarp();
  return R;
}

// Complexity left behind by a macroblock for the row below (out_top) and the MB to the right
// (out_left): encode_residual_data / record_residual_stats bookkeeping + the skip clear
// (vp8.rs:1367-1374, :1468-1474, Complexity::clear :138).
__device__ __forceinline__ void complexity_after(bool is_b, bool skip, int y2nz, u32 ynz, u32 uvnz, u32 in_top, u32 in_left,
                                                 u32& out_top, u32& out_left) {
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

// ---------------------------------------------------------------------------------------------
// Pass-1 chroma chain: one warp per image, raster order (see file header).  A latency-bound
// chain (1024 warps for 1024 images).  It reads nothing luma produces: it writes the chroma levels
// of rec1, the chroma borders, derr1 and c1info; k_finish1 then derives the skip flags and the
// complexity contexts from both halves.  (Measured: running it UNDER the luma wavefront, as extra
// tickets of k_search<1> or on a side stream, starves the chains -- instruction-cache contention --
// and is slower than running the two kernels back to back.)
// ---------------------------------------------------------------------------------------------
__device__ void chroma_chain1(WarpScratch& W, const ChunkParams& P, u32 img, int lane) {
  const ImageDesc d = P.img[img];
  const ImageState& IS = P.st[img];
  const int mbw = d.mbw, mbh = d.mbh, pw = mbw * 16, cwid = mbw * 8;
  const u32 nmb = (u32)mbw * mbh;
  const u8* yp = P.planes + d.y_off;
  const u8* up = yp + (size_t)pw * mbh * 16;
  const u8* vp = up + (size_t)cwid * mbh * 8;
  CostCtx cc;
  cc.probs = ZW_TAB(kCoeffProbs);
  cc.level_cost = nullptr;
  const bool seg_on = IS.seg_enabled != 0;
  u32 left_derr = 0;  // carried across rows in pass 1 (Q5)
  // Everything a macroblock needs from global memory except its left neighbour (source rows, the
  // bottom row / diffusion state of the macroblock above, its segment) is known one step ahead: it is
  // fetched while the previous macroblock is processed -- the chain is latency bound and these were
  // three exposed L2 round trips per macroblock.  Not possible when the macroblock above IS the previous
  // one (mbw == 1).
  const bool ahead = mbw > 1;
  uint2 n_src = make_uint2(0, 0);
  u32 n_top = 127, n_derr = 0, n_seg = 0;
  // the (at most four) quantiser sets of the image, copied next to the scratch once: a macroblock's parameters are then a
  // shared-memory read away instead of two dependent global loads at the head of every step of the chain
  SegParams* seg4 = reinterpret_cast<SegParams*>(W.coef);  // pass-1 chroma never touches the luma trellis buffer
  static_assert(4 * sizeof(SegParams) <= sizeof(W.coef), "segment parameters must fit the borrowed buffer");
  for (int k = lane; k < (int)(4 * sizeof(SegParams) / 4); k += 32) {
    const int sgm = k / (int)(sizeof(SegParams) / 4), wd = k % (int)(sizeof(SegParams) / 4);
    const u32 qi = seg_on ? (u32)IS.seg_qidx[sgm] : P.base_qidx;
    reinterpret_cast<u32*>(seg4)[k] = reinterpret_cast<const u32*>(&P.segtab[qi])[wd];
  }
  __syncwarp();
  // Every value is ONE load per lane whose result is not touched until the next step (addresses are selected, not values:
  // a select or a conversion right after the load would make the warp wait for the L2 round trip it is meant to hide --
  // measured: 600 cycles per macroblock).  Lanes 0..15: source row (8 bytes) and one pixel of the row above; every lane:
  // the diffusion state of the macroblock above and the segment.
  const u8* src_row = lane < 8 ? up + (size_t)(lane & 7) * cwid : vp + (size_t)(lane & 7) * cwid;
  auto fetch = [&](u32 i, int fx, int fy, uint2& o_src, u32& o_top, u32& o_derr, u32& o_seg) {
    const u32 g = d.mb_off + i;
    if (lane < 16) o_src = __ldg(reinterpret_cast<const uint2*>(src_row + (size_t)(fy * 8) * cwid + fx * 8));
    if (fy > 0) {
      const MbBottom* bt = &P.bottom[g - mbw];
      const u8* tp = lane < 8 ? &bt->u[lane & 7] : &bt->v[lane & 7];
      if (lane < 16) o_top = __ldcg(tp);
      o_derr = __ldcg(&P.derr1[g - mbw]);
    } else {
      o_top = 127; o_derr = 0;
    }
    if (seg_on) o_seg = (u32)P.segmap[g];
  };
  if (ahead) fetch(0, 0, 0, n_src, n_top, n_derr, n_seg);
  int mbx = 0, mby = 0;
  u32 lcap_u = 129, lcap_v = 129;
  for (u32 i = 0; i < nmb; i++) {
    const u32 gmb = d.mb_off + i;
    uint2 c_src = n_src;
    u32 c_top = n_top, top_derr = n_derr, c_seg = n_seg;
    const int nx = mbx + 1 == mbw ? 0 : mbx + 1, ny = mbx + 1 == mbw ? mby + 1 : mby;  // the next macroblock in raster order
    if (!ahead) fetch(i, mbx, mby, c_src, c_top, top_derr, c_seg);
    const SegParams& SP = seg4[seg_on ? c_seg : 0u];
    // stage the macroblock (load_chroma_mb with the prefetched values)
    if (lane < 8) *reinterpret_cast<uint2*>(&W.src_u[lane * 8]) = c_src;
    else if (lane < 16) *reinterpret_cast<uint2*>(&W.src_v[(lane - 8) * 8]) = c_src;
    // top row, left column and corner in ONE phase: the left column / corner were captured in registers from the previous
    // macroblock's reconstruction (lane l < 8: row 1 + l, lane 8: the corner), so nothing is read here
    if (mby == 0) {
      if (lane >= 1 && lane != 16) W.uvws[lane] = 127;  // [0] and [16] are the corners, written below
    } else {
      if (lane < 8) W.uvws[1 + lane] = (u8)c_top;
      else if (lane < 16) W.uvws[17 + (lane - 8)] = (u8)c_top;
    }
    if (lane < 8) {
      W.uvws[(1 + lane) * 32] = (mbx == 0) ? 129 : (u8)lcap_u;
      W.uvws[(1 + lane) * 32 + 16] = (mbx == 0) ? 129 : (u8)lcap_v;
    }
    if (lane == 8) {
      W.uvws[0] = (mby == 0) ? 127 : (mbx == 0 ? 129 : (u8)lcap_u);
      W.uvws[16] = (mby == 0) ? 127 : (mbx == 0 ? 129 : (u8)lcap_v);
    }
    __syncwarp();
    // the next macroblock's loads are issued only now: a __syncwarp orders memory, i.e. waits for the loads in flight, and
    // the next one is a whole mode search away (issued before the two above, the prefetch stalled the chain for an L2 round trip)
    if (ahead && i + 1 < nmb) fetch(i + 1, nx, ny, n_src, n_top, n_derr, n_seg);
    const ChromaOut C = chroma_mb(W, SP, cc, mbx, mby, left_derr, top_derr, lane);
    MbRecord* r = &P.rec1[gmb];
    if (lane < 8) {
      u32* g = reinterpret_cast<u32*>(r->levels[17 + lane]);
      const u32* sl = reinterpret_cast<const u32*>(W.rec.levels[17 + lane]);
#pragma unroll
      for (int k = 0; k < 8; k++) g[k] = sl[k];
    }
    if (lane == 0) {
      P.derr1[gmb] = top_derr;
      P.c1info[2 * gmb] = left_derr;
      P.c1info[2 * gmb + 1] = (u32)C.uv_mode | (C.uvnz << 8);
    }
    // left column (lanes 0..7: rows 1..8) and corner (lane 8: row 0 = the row above, untouched by the reconstruction) of the next macroblock
    if (lane < 9) { const int rr = lane < 8 ? 1 + lane : 0; lcap_u = W.uvws[rr * 32 + 8]; lcap_v = W.uvws[rr * 32 + 24]; }
    MbBottom* bo = &P.bottom[gmb];
    if (lane >= 16 && lane < 24) bo->u[lane - 16] = W.uvws[8 * 32 + 1 + (lane - 16)];
    if (lane >= 24) bo->v[lane - 24] = W.uvws[8 * 32 + 17 + (lane - 24)];
    __syncwarp();
    mbx = nx; mby = ny;
  }
}

// ---------------------------------------------------------------------------------------------
// The wavefront kernel.  PASS 1: luma only (see file header).  PASS 2: luma + chroma.
// ---------------------------------------------------------------------------------------------
#ifndef ZW_LS_PASS1
#define ZW_LS_PASS1 0  // measured: pass 1 (9-14 % "no instruction" stalls) does not gain from the lock step
#endif
__host__ __device__ constexpr bool search_lockstep(int pass) { return (pass == 2 || ZW_LS_PASS1) && LS_WARPS > 0; }
__host__ __device__ constexpr int search_warps(int pass) { return search_lockstep(pass) ? LS_WARPS : SEARCH_WARPS; }
__host__ __device__ constexpr int search_min_blocks(int pass) {
#ifdef ZW_LS_MIN_BLOCKS
  return search_lockstep(pass) ? ZW_LS_MIN_BLOCKS : ZW_SEARCH_MIN_BLOCKS;
#else
  return search_lockstep(pass) ? 20 / (LS_WARPS > 0 ? LS_WARPS : 1) : ZW_SEARCH_MIN_BLOCKS;  // lock-step CTAs: 20 warps per SM (96 registers)
#endif
}
// dynamic shared memory of a wavefront kernel launched with `nwarps` warps per CTA
__host__ __device__ constexpr size_t search_smem_bytes(int nwarps) { return sizeof(SearchShared) + (size_t)(nwarps - SEARCH_WARPS) * sizeof(WarpScratch); }
// k_search<2> keeps the I4 (type 3) level costs [8][3][68] of each warp's image behind the scratch: the most-read table of
// the pass (measured: pass 2 39.0 -> 38.6 ms at method 4, 54.8 -> 52.5 ms at method 6; L1 alone thrashes with 20 images per SM)
constexpr int LC3_ENTRIES = 8 * 3 * 68;
__host__ __device__ constexpr size_t search_smem_total(int pass) {
  return search_smem_bytes(search_warps(pass)) + (pass == 2 ? (size_t)search_warps(pass) * LC3_ENTRIES * sizeof(u16) : 0);
}

template <int PASS>
__global__ void __launch_bounds__(search_warps(PASS) * 32, search_min_blocks(PASS)) k_search(ChunkParams P) {
  constexpr bool LS = search_lockstep(PASS);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SearchShared& SH = *reinterpret_cast<SearchShared*>(smem_raw);
  __shared__ int s_active;  // LS: warps of this CTA that still have (or may get) a row
  // (mode, pixel) -> dtab slot; the TM row (flat entries 16..31) points at dtab[32 + n], where the TM pixels live -- one
  // writer per entry
  for (int i = threadIdx.x; i < 160; i += blockDim.x) (&SH.pred_idx[0][0])[i] = (i >= 16 && i < 32) ? (u8)(16 + i) : (&d_pred_idx[0][0])[i];
  if (threadIdx.x < 32) { SH.dtaps[threadIdx.x] = d_dtaps[threadIdx.x]; fill_lane_consts(SH.lk, threadIdx.x); fill_lane_consts8(SH.lk8, threadIdx.x); }
  if (threadIdx.x == 0) s_active = (int)(blockDim.x >> 5);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  WarpScratch& W = SH.w[threadIdx.x >> 5];
  int* progress = P.progress + (PASS - 1) * P.n_rows;
  MbRecord* recs = PASS == 1 ? P.rec1 : P.rec2;
  const bool trellis = (PASS == 2) && P.do_trellis;

  // state of the row this warp owns
  bool have_row = false, done = false;
  u32 img = 0, row_mb0 = 0, up_mb0 = 0, left_nz = 0, mb_off = 0, row_off = 0;
  int mbw = 0, mby = 0, mbx = 0, seen = 0;
  const u8* yp = nullptr;
  bool seg_on = false;
  CostCtx cc;
  cc.probs = nullptr; cc.level_cost = nullptr;
#ifdef ZW_WAIT_STATS
  long long r0 = 0;
#endif

  for (;;) {
    if (!have_row && !done) {
      u32 t = 0;
      if (lane == 0) t = atomicAdd(&P.ticket[PASS - 1], 1u);
      t = __shfl_sync(FULL, t, 0);
      if (t >= P.n_rows) {
        if (!LS) break;
        done = true;
        if (lane == 0) atomicSub(&s_active, 1);
      } else {
#ifdef ZW_WAIT_STATS
        r0 = clock64();
#endif
        const RowRef rr = P.rows[t];
        const ImageDesc d = P.img[rr.img];
        img = rr.img;
        mbw = d.mbw; mby = rr.mby; mb_off = d.mb_off; row_off = d.row_off;
        yp = P.planes + d.y_off;
        row_mb0 = d.mb_off + mby * mbw;
        up_mb0 = row_mb0 - mbw;  // only dereferenced when mby > 0
        cc.probs = PASS == 1 ? ZW_TAB(kCoeffProbs) : P.probs + (size_t)img * 1056;
        cc.level_cost = PASS == 1 ? nullptr : P.lcost + (size_t)img * 6528;
        cc.lc3 = nullptr;
        if (PASS == 2) {  // the image's I4 level costs (3.2 KB): the most-read table of the pass, kept in shared memory per row
          const u32* src3 = reinterpret_cast<const u32*>(cc.level_cost + 3 * 1632);
          u16* lc3 = reinterpret_cast<u16*>(smem_raw + search_smem_bytes(search_warps(PASS))) + (threadIdx.x >> 5) * LC3_ENTRIES;
          for (int k = lane; k < LC3_ENTRIES / 2; k += 32) reinterpret_cast<u32*>(lc3)[k] = __ldg(src3 + k);
          cc.lc3 = lc3;
        }
        seg_on = P.st[img].seg_enabled != 0;
        // row-start state (vp8.rs:1339-1344 / :1423-1429)
        left_nz = 0; mbx = 0; seen = 0;
        if (lane < 17) W.left_y[lane] = 129;
        {  // the image's I4 (type 3) end-of-block / first-branch costs, see coop_cost_i4
          const u8* pr3 = cc.probs + 3 * 264;
          if (lane < 16) {
            const int nb = SH.lk[6][lane].y;  // band(n + 1)
            W.eob_pack[lane] = lane < 15 ? (bit_cost(0, pr3[(nb * 3 + 1) * 11]) | (bit_cost(0, pr3[(nb * 3 + 2) * 11]) << 16)) : 0u;
          } else if (lane < 20) {
            const int k = lane - 16;
            W.p0c[k] = k < 3 ? bit_cost(0, pr3[k * 11]) : bit_cost(1, pr3[0]);
          }
          if (PASS == 2) {  // I16 AC blocks (type 0, first 1 -> band 1) for the trellis
            const u8* pr0 = cc.probs;
            if (lane < 16) {
              const int nb = SH.lk[6][lane].y;
              W.eob_pack0[lane] = lane < 15 ? (bit_cost(0, pr0[(nb * 3 + 1) * 11]) | (bit_cost(0, pr0[(nb * 3 + 2) * 11]) << 16)) : 0u;
            } else if (lane < 20) {
              const int k = lane - 16;
              W.p0c0[k] = k < 3 ? bit_cost(0, pr0[(3 + k) * 11]) : bit_cost(1, pr0[3 * 11]);
            }
          }
        }
        __syncwarp();
        have_row = true;
      }
    }
    // ---- dependency: the top / top-right neighbours (row above finished mbx + 1) ----
    bool work = have_row;
    if (work && mby > 0) {
      // start_slack > 2 makes a row START only once the row above is that many macroblocks ahead
      // (tuning knob, default off: measured with ZW_WAIT_STATS, rows wait ~1 % of their time).
      const int need = min(mbx == 0 ? max(2, (int)P.start_slack) : mbx + 2, mbw);
      if (LS) {  // never block inside a lock-step round: look once, sit the round out if not ready
        int ok = 1;
        if (lane == 0 && seen < need) {
          seen = ld_flag(&progress[row_off + mby - 1]);
          ok = seen >= need;
          if (ok) fence_acquire();
        }
        work = __shfl_sync(FULL, ok, 0) != 0;
      } else {
        if (lane == 0 && seen < need) {
#ifdef ZW_WAIT_STATS
          const long long w0 = clock64();
#endif
          while ((seen = ld_flag(&progress[row_off + mby - 1])) < need) __nanosleep(100);
          fence_acquire();
#ifdef ZW_WAIT_STATS
          atomicAdd(reinterpret_cast<unsigned long long*>(P.ticket) + 2 + PASS, (unsigned long long)(clock64() - w0));
#endif
        }
        __syncwarp();
      }
    }
    if (LS) {
      __syncthreads();            // phase barrier 1 (s_active only changes at ticket time, i.e. before it)
      if (s_active == 0) break;   // every warp reads the same value: the next change comes after barrier 3
    }
    const u32 gmb = row_mb0 + mbx;
    int seg = 0;
    u32 top_nz = 0;
    const SegParams* SPp = &P.segtab[P.base_qidx];
    if (work) {
      seg = seg_on ? P.segmap[gmb] : 0;
      SPp = &P.segtab[seg_on ? P.st[img].seg_qidx[seg] : P.base_qidx];
      if (PASS == 2 && mby > 0) top_nz = __ldcg(&P.nz_after[up_mb0 + mbx]);
      load_luma_mb(W, P, yp, mbw * 16, mbw, mbx, mby, up_mb0, lane);
      for (int k = lane; k < 136; k += 32) reinterpret_cast<u32*>(W.rec.levels)[k] = 0;  // luma levels [0..16]
      __syncwarp();
    }
    const SegParams& SP = *SPp;
    const LumaOut L = luma_mb<LS, (PASS == 1 ? (ZW_I4_LANES8 & 1) : (ZW_I4_LANES8 & 2)) != 0>(W, SH, SP, cc, (int)P.i4_modes, P.i4_always != 0, trellis, mbx, mby, top_nz, left_nz, lane, work);
    if (!work) continue;
    bool skip = false;
    u32 out_top = 0, out_left = 0;
    if (PASS == 2) {
      // chroma of this macroblock was coded by k_chroma2 (it does not depend on luma)
      const u32 uvnz = P.uvflags[gmb];
      skip = !(L.simple_nz || uvnz != 0);
      complexity_after(L.use_i4, skip, L.y2nz, L.ynz, uvnz, top_nz, left_nz, out_top, out_left);
    }
    if (lane == 0) {
      W.rec.ymode = L.use_i4 ? 4 : (u8)L.mode16;
      W.rec.segment = (u8)seg;
      W.rec.skip = skip;
      if (PASS == 2) {
        W.rec.top_nz = (u16)top_nz;
        W.rec.left_nz = (u16)left_nz;
        // chroma-owned header fields: keep what k_chroma2 stored
        W.rec.uvmode = recs[gmb].uvmode;
        *reinterpret_cast<u32*>(W.rec.derr_left) = *reinterpret_cast<const u32*>(recs[gmb].derr_left);
        *reinterpret_cast<u32*>(W.rec.derr_top) = *reinterpret_cast<const u32*>(recs[gmb].derr_top);
      } else {
        // pass 1: luma flags parked here until k_finish1 completes the record
        W.rec.uvmode = 0;
        W.rec.top_nz = (u16)L.ynz;
        W.rec.left_nz = (u16)((L.y2nz ? 1 : 0) | (L.simple_nz ? 2 : 0));
        *reinterpret_cast<u32*>(W.rec.derr_left) = 0;
        *reinterpret_cast<u32*>(W.rec.derr_top) = 0;
      }
    }
    if (lane < 16) W.rec.bmodes[lane] = L.use_i4 ? W.bmodes[lane] : 0;
    __syncwarp();
    {
      // header (8 words) + luma levels (136 words); a skipped MB codes nothing: zero all 200 level words
      const u32* sr = reinterpret_cast<const u32*>(&W.rec);
      u32* g = reinterpret_cast<u32*>(&recs[gmb]);
      if (PASS == 2 && skip) {
        for (int k = lane; k < 208; k += 32) g[k] = k < 8 ? sr[k] : 0u;
      } else {
        for (int k = lane; k < 144; k += 32) g[k] = sr[k];  // chroma levels: k_chroma2 / the pass-1 chroma chain
      }
    }
    left_nz = out_left;
    // borders for the neighbours
    if (lane < 17) W.left_y[lane] = W.yws[lane * 32 + 16];
    MbBottom* bo = &P.bottom[gmb];
    if (lane < 16) bo->y[lane] = W.yws[16 * 32 + 1 + lane];
    if (PASS == 2 && lane == 0) P.nz_after[gmb] = (u16)out_top;
    __syncwarp();
    if (lane == 0) publish_flag(&progress[row_off + mby], mbx + 1);
    mbx++;
    if (mbx == mbw) {
      have_row = false;
#ifdef ZW_WAIT_STATS
      if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(P.ticket) + 4 + PASS, (unsigned long long)(clock64() - r0));
#endif
    }
  }
  (void)mb_off;
}

#ifndef ZW_CHROMA1_MIN_BLOCKS
#define ZW_CHROMA1_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(SEARCH_WARPS * 32, ZW_CHROMA1_MIN_BLOCKS) k_chroma1(ChunkParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SearchShared& SH = *reinterpret_cast<SearchShared*>(smem_raw);
  const int lane = threadIdx.x & 31;
  WarpScratch& W = SH.w[threadIdx.x >> 5];
  for (;;) {
    u32 img = 0;
    if (lane == 0) img = atomicAdd(&P.ticket[2], 1u);
    img = __shfl_sync(FULL, img, 0);
    if (img >= P.n_img) break;
    chroma_chain1(W, P, img, lane);
  }
}

// Completes rec1 after k_search<1>: skip flags, incoming complexity contexts (the bookkeeping of
// encode_residual_data / record_residual_stats, vp8.rs:1367-1374, :1468-1474), chroma header
// fields, zeroed levels of skipped macroblocks, the pass-1 skip count.  One CTA per image:
//   A  per MB: skip + the complexity it leaves behind, Y2 bit of I4 macroblocks left open
//   B  the Y2 bit passes through I4 macroblocks: one thread per column (top) / per row (left)
//   C  per MB: contexts = what the neighbours left behind; finish the record
// Scratch: two u16 per MB in derr2 (free until k_chroma2), bit 15 = is_b, bit 14 = skip.
__global__ void __launch_bounds__(256) k_finish1(ChunkParams P) {
  const u32 img = blockIdx.x;
  const ImageDesc d = P.img[img];
  const u32 mbw = d.mbw, mbh = d.mbh, nmb = mbw * mbh;
  u16* sc = reinterpret_cast<u16*>(P.derr2) + 2 * (size_t)d.mb_off;
  MbRecord* recs = P.rec1 + d.mb_off;
  const u32* ci = P.c1info + 2 * (size_t)d.mb_off;
  __shared__ u32 s_nskip;
  if (threadIdx.x == 0) s_nskip = 0;
  for (u32 i = threadIdx.x; i < nmb; i += blockDim.x) {
    const MbRecord& r = recs[i];
    const u32 ynz = r.top_nz, lf = r.left_nz;  // parked by k_search<1>
    const bool is_b = r.ymode == 4;
    const u32 uvnz = (ci[2 * i + 1] >> 8) & 0xffu;
    const bool skip = !((lf & 2) || uvnz != 0);
    u32 ot, ol;
    complexity_after(is_b, skip, (int)(lf & 1), ynz, uvnz, 0, 0, ot, ol);
    const u32 fl = (is_b ? 0x8000u : 0u) | (skip ? 0x4000u : 0u);
    sc[2 * i] = (u16)(ot | fl);
    sc[2 * i + 1] = (u16)(ol | fl);
  }
  __syncthreads();
  for (u32 j = threadIdx.x; j < mbw + mbh; j += blockDim.x) {
    u32 y2 = 0;
    if (j < mbw) {
      for (u32 y = 0; y < mbh; y++) {
        const u32 k = 2 * (y * mbw + j);
        const u32 v = sc[k];
        if (v & 0x8000u) sc[k] = (u16)((v & ~1u) | y2); else y2 = v & 1;
      }
    } else {
      const u32 y = j - mbw;
      for (u32 x = 0; x < mbw; x++) {
        const u32 k = 2 * (y * mbw + x) + 1;
        const u32 v = sc[k];
        if (v & 0x8000u) sc[k] = (u16)((v & ~1u) | y2); else y2 = v & 1;
      }
    }
  }
  __syncthreads();
  u32 nskip = 0;
  for (u32 i = threadIdx.x; i < nmb; i += blockDim.x) {
    const u32 x = i % mbw, y = i / mbw;
    MbRecord& r = recs[i];
    const bool skip = (sc[2 * i] & 0x4000u) != 0;
    r.top_nz = (u16)(y > 0 ? (sc[2 * (i - mbw)] & 0x1ffu) : 0u);
    r.left_nz = (u16)(x > 0 ? (sc[2 * (i - 1) + 1] & 0x1ffu) : 0u);
    r.skip = skip;
    r.uvmode = (u8)(ci[2 * i + 1] & 0xffu);
    *reinterpret_cast<u32*>(r.derr_left) = ci[2 * i];
    *reinterpret_cast<u32*>(r.derr_top) = P.derr1[d.mb_off + i];
    if (skip) {
      uint4* lv = reinterpret_cast<uint4*>(r.levels);
      for (int k = 0; k < 50; k++) lv[k] = make_uint4(0, 0, 0, 0);
    }
    nskip += skip;
  }
  atomicAdd(&s_nskip, nskip);
  __syncthreads();
  if (threadIdx.x == 0) P.st[img].n_skip1 = s_nskip;
}

// ---------------------------------------------------------------------------------------------
// Pass-2 chroma: wavefront over rows (pass 2 resets left_derr per row, so rows only depend on
// the row above).  Runs BEFORE k_search<2>: chroma never depends on luma, while the luma kernel
// needs the chroma non-zero flags for the skip decision and the complexity contexts.  Keeping
// chroma out of the luma kernel also keeps both instruction working sets small.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEARCH_WARPS * 32, ZW_SEARCH_MIN_BLOCKS) k_chroma2(ChunkParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SearchShared& SH = *reinterpret_cast<SearchShared*>(smem_raw);
  const int lane = threadIdx.x & 31;
  WarpScratch& W = SH.w[threadIdx.x >> 5];
  int* progress = P.progress + 2 * P.n_rows;
  for (;;) {
    u32 t = 0;
    if (lane == 0) t = atomicAdd(&P.ticket[3], 1u);
    t = __shfl_sync(FULL, t, 0);
    if (t >= P.n_rows) break;
    const RowRef rr = P.rows[t];
    const ImageDesc d = P.img[rr.img];
    const ImageState& IS = P.st[rr.img];
    const int mbw = d.mbw, mby = rr.mby, pw = mbw * 16, cwid = mbw * 8;
    const u8* yp = P.planes + d.y_off;
    const u8* up = yp + (size_t)pw * d.mbh * 16;
    const u8* vp = up + (size_t)cwid * d.mbh * 8;
    const u32 row_mb0 = d.mb_off + mby * mbw, up_mb0 = row_mb0 - mbw;
    CostCtx cc;
    cc.probs = P.probs + (size_t)rr.img * 1056;
    cc.level_cost = P.lcost + (size_t)rr.img * 6528;
    const bool seg_on = IS.seg_enabled != 0;
    u32 left_derr = 0;  // reset per row in pass 2 (vp8.rs:1425)
    if (lane < 9) { W.left_u[lane] = 129; W.left_v[lane] = 129; }
    __syncwarp();
    int seen = 0;
    for (int mbx = 0; mbx < mbw; mbx++) {
      if (mby > 0) {  // chroma needs the macroblock above only (no top-right)
        const int need = min(mbx == 0 ? max(1, (int)P.start_slack) : mbx + 1, mbw);
        if (lane == 0 && seen < need) {
          while ((seen = ld_flag(&progress[d.row_off + mby - 1])) < need) __nanosleep(100);
          fence_acquire();
        }
        __syncwarp();
      }
      const u32 gmb = row_mb0 + mbx;
      const int seg = seg_on ? P.segmap[gmb] : 0;
      const SegParams& SP = P.segtab[seg_on ? IS.seg_qidx[seg] : P.base_qidx];
      // Q4: top_derr is not reset between passes -> row 0 of pass 2 starts from pass 1's last row
      u32 top_derr = mby > 0 ? __ldcg(&P.derr2[up_mb0 + mbx]) : __ldcg(&P.derr1[d.mb_off + (d.mbh - 1) * mbw + mbx]);
      load_chroma_mb(W, P, up, vp, cwid, mbx, mby, up_mb0, lane);
      const ChromaOut C = chroma_mb(W, SP, cc, mbx, mby, left_derr, top_derr, lane);
      MbRecord* r = &P.rec2[gmb];
      if (lane < 8) {
        u32* g = reinterpret_cast<u32*>(r->levels[17 + lane]);
        const u32* s = reinterpret_cast<const u32*>(W.rec.levels[17 + lane]);
#pragma unroll
        for (int k = 0; k < 8; k++) g[k] = s[k];
      }
      if (lane == 0) {
        r->uvmode = (u8)C.uv_mode;
        *reinterpret_cast<u32*>(r->derr_left) = left_derr;
        *reinterpret_cast<u32*>(r->derr_top) = top_derr;
        P.derr2[gmb] = top_derr;
        P.uvflags[gmb] = (u8)C.uvnz;
      }
      if (lane < 9) { W.left_u[lane] = W.uvws[lane * 32 + 8]; W.left_v[lane] = W.uvws[lane * 32 + 24]; }
      MbBottom* bo = &P.bottom[gmb];
      if (lane >= 16 && lane < 24) bo->u[lane - 16] = W.uvws[8 * 32 + 1 + (lane - 16)];
      if (lane >= 24) bo->v[lane - 24] = W.uvws[8 * 32 + 17 + (lane - 24)];
      __syncwarp();
      if (lane == 0) publish_flag(&progress[d.row_off + mby], mbx + 1);
    }
  }
}

}  // namespace zw
#endif
