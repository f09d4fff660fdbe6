// zw_prims.cuh -- per-lane integer primitives of the VP8 encoder core (sm_100a).
//
// Everything here is straight-line 32-bit integer code on register arrays: one lane owns one
// 4x4 block.  The functions are ZW_HD (__host__ __device__) only so that tests/ can compile this
// header with g++ and compare each primitive with the CPU oracle on random inputs without a
// GPU; the shipped library only ever calls them from kernels.
//
// Reference semantics (file:line under /root/reference):
//   fdct4x4   src/common/transform.rs:176-207   (libwebp FTransform rounding)
//   idct4x4   src/common/transform.rs:35-79
//   wht4x4    src/common/transform.rs:116-158,  iwht4x4 :82-114
//   quantdiv  src/encoder/cost.rs:244,  quantize_coeff :457, dequantize :484
//   t_transform src/encoder/cost.rs:73-107
// 32-bit safety: every reachable intermediate fits i32 (|residual| <= 255, |coeff| < 2^15,
// iq <= 2^15, see DESIGN.md "integer ranges"); the reference widens to i64 only defensively.
#ifndef ZW_PRIMS_CUH
#define ZW_PRIMS_CUH
#include <stdint.h>

#if defined(__CUDACC__)
#define ZW_HD __host__ __device__ __forceinline__
#define ZW_D __device__ __forceinline__
#else
#define ZW_HD inline
#define ZW_D inline
#endif

namespace zw {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int8_t i8;
typedef int16_t i16;
typedef int32_t i32;
typedef int64_t i64;

ZW_HD i32 iabs(i32 v) { return v < 0 ? -v : v; }
ZW_HD i32 imin(i32 a, i32 b) { return a < b ? a : b; }
ZW_HD i32 imax(i32 a, i32 b) { return a > b ? a : b; }
ZW_HD i32 clip255(i32 v) { return imin(imax(v, 0), 255); }

// Forward 4x4 DCT in place, natural (row-major) coefficient order.
ZW_HD void fdct4x4(i32* b) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a = (b[i * 4] + b[i * 4 + 3]) * 8;
    i32 bb = (b[i * 4 + 1] + b[i * 4 + 2]) * 8;
    i32 c = (b[i * 4 + 1] - b[i * 4 + 2]) * 8;
    i32 d = (b[i * 4] - b[i * 4 + 3]) * 8;
    b[i * 4] = a + bb;
    b[i * 4 + 2] = a - bb;
    b[i * 4 + 1] = (c * 2217 + d * 5352 + 14500) >> 12;
    b[i * 4 + 3] = (d * 2217 - c * 5352 + 7500) >> 12;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a = b[i] + b[i + 12];
    i32 bb = b[i + 4] + b[i + 8];
    i32 c = b[i + 4] - b[i + 8];
    i32 d = b[i] - b[i + 12];
    b[i] = (a + bb + 7) >> 4;
    b[i + 8] = (a - bb + 7) >> 4;
    b[i + 4] = ((c * 2217 + d * 5352 + 12000) >> 16) + (d != 0 ? 1 : 0);
    b[i + 12] = (d * 2217 - c * 5352 + 51000) >> 16;
  }
}

// Inverse 4x4 DCT in place.
ZW_HD void idct4x4(i32* b) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a1 = b[i] + b[8 + i];
    i32 b1 = b[i] - b[8 + i];
    i32 t1 = (b[4 + i] * 35468) >> 16;
    i32 t2 = b[12 + i] + ((b[12 + i] * 20091) >> 16);
    i32 c1 = t1 - t2;
    t1 = b[4 + i] + ((b[4 + i] * 20091) >> 16);
    t2 = (b[12 + i] * 35468) >> 16;
    i32 d1 = t1 + t2;
    b[i] = a1 + d1;
    b[4 + i] = b1 + c1;
    b[12 + i] = a1 - d1;
    b[8 + i] = b1 - c1;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a1 = b[4 * i] + b[4 * i + 2];
    i32 b1 = b[4 * i] - b[4 * i + 2];
    i32 t1 = (b[4 * i + 1] * 35468) >> 16;
    i32 t2 = b[4 * i + 3] + ((b[4 * i + 3] * 20091) >> 16);
    i32 c1 = t1 - t2;
    t1 = b[4 * i + 1] + ((b[4 * i + 1] * 20091) >> 16);
    t2 = (b[4 * i + 3] * 35468) >> 16;
    i32 d1 = t1 + t2;
    b[4 * i] = (a1 + d1 + 4) >> 3;
    b[4 * i + 3] = (a1 - d1 + 4) >> 3;
    b[4 * i + 1] = (b1 + c1 + 4) >> 3;
    b[4 * i + 2] = (b1 - c1 + 4) >> 3;
  }
}

ZW_HD void wht4x4(i32* b) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a = b[i * 4] + b[i * 4 + 3];
    i32 bb = b[i * 4 + 1] + b[i * 4 + 2];
    i32 c = b[i * 4 + 1] - b[i * 4 + 2];
    i32 d = b[i * 4] - b[i * 4 + 3];
    b[i * 4] = a + bb;
    b[i * 4 + 1] = c + d;
    b[i * 4 + 2] = a - bb;
    b[i * 4 + 3] = d - c;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a1 = b[i] + b[i + 12];
    i32 b1 = b[i + 4] + b[i + 8];
    i32 c1 = b[i + 4] - b[i + 8];
    i32 d1 = b[i] - b[i + 12];
    i32 a2 = a1 + b1, b2 = c1 + d1, c2 = a1 - b1, d2 = d1 - c1;
    // (x + (x > 0)) / 2 with truncation toward zero
    b[i] = (a2 + (a2 > 0 ? 1 : 0)) / 2;
    b[i + 4] = (b2 + (b2 > 0 ? 1 : 0)) / 2;
    b[i + 8] = (c2 + (c2 > 0 ? 1 : 0)) / 2;
    b[i + 12] = (d2 + (d2 > 0 ? 1 : 0)) / 2;
  }
}

ZW_HD void iwht4x4(i32* b) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a1 = b[i] + b[12 + i];
    i32 b1 = b[4 + i] + b[8 + i];
    i32 c1 = b[4 + i] - b[8 + i];
    i32 d1 = b[i] - b[12 + i];
    b[i] = a1 + b1;
    b[4 + i] = c1 + d1;
    b[8 + i] = a1 - b1;
    b[12 + i] = d1 - c1;
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    i32 a1 = b[4 * r] + b[4 * r + 3];
    i32 b1 = b[4 * r + 1] + b[4 * r + 2];
    i32 c1 = b[4 * r + 1] - b[4 * r + 2];
    i32 d1 = b[4 * r] - b[4 * r + 3];
    i32 a2 = a1 + b1, b2 = c1 + d1, c2 = a1 - b1, d2 = d1 - c1;
    b[4 * r] = (a2 + 3) >> 3;
    b[4 * r + 1] = (b2 + 3) >> 3;
    b[4 * r + 2] = (c2 + 3) >> 3;
    b[4 * r + 3] = (d2 + 3) >> 3;
  }
}

// Quantiser for one coefficient type: position 0 uses the DC entry, 1..15 the AC entry
// (VP8Matrix::new replicates the AC values, cost.rs:430-436).
struct Matrix {
  u16 q[2];
  u32 iq[2];
  u32 bias[2];
};

ZW_HD i32 quantdiv(u32 coeff, u32 iq, u32 bias) { return (i32)((coeff * iq + bias) >> 17); }

ZW_HD i32 quantize_coeff(i32 coeff, const Matrix& m, int pos) {
  int k = pos > 0;
  i32 a = iabs(coeff);
  i32 level = quantdiv((u32)a, m.iq[k], m.bias[k]);
  return coeff < 0 ? -level : level;
}
ZW_HD i32 dequantize(i32 level, const Matrix& m, int pos) { return level * (i32)m.q[pos > 0]; }

// Weighted 4x4 Hadamard of a pixel block given as 16 values (row-major).
ZW_HD i32 t_transform16(const i32* in, const u16* w) {
  i32 tmp[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a0 = in[i * 4] + in[i * 4 + 2];
    i32 a1 = in[i * 4 + 1] + in[i * 4 + 3];
    i32 a2 = in[i * 4 + 1] - in[i * 4 + 3];
    i32 a3 = in[i * 4] - in[i * 4 + 2];
    tmp[i * 4] = a0 + a1;
    tmp[i * 4 + 1] = a3 + a2;
    tmp[i * 4 + 2] = a3 - a2;
    tmp[i * 4 + 3] = a0 - a1;
  }
  i32 sum = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    i32 a0 = tmp[i] + tmp[8 + i];
    i32 a1 = tmp[4 + i] + tmp[12 + i];
    i32 a2 = tmp[4 + i] - tmp[12 + i];
    i32 a3 = tmp[i] - tmp[8 + i];
    i32 b0 = a0 + a1, b1 = a3 + a2, b2 = a3 - a2, b3 = a0 - a1;
    sum += (i32)w[i] * iabs(b0);
    sum += (i32)w[4 + i] * iabs(b1);
    sum += (i32)w[8 + i] * iabs(b2);
    sum += (i32)w[12 + i] * iabs(b3);
  }
  return sum;
}

// The ten 4x4 intra predictors from the 13 edge pixels (src/common/prediction.rs:326-780).
// e[0..3] = l3 l2 l1 l0, e[4] = p (top-left), e[5..12] = a0..a7.
// Modes: 0 DC, 1 TM, 2 VE, 3 HE, 4 LD, 5 RD, 6 VR, 7 VL, 8 HD, 9 HU.
// Modes 2..9 are pure 3-tap filters (avg3(x,y,z) = (x+2y+z+2)>>2; avg2(x,y) == avg3(x,y,x) ...
// careful: avg2(x,y) = (x+y+1)>>1 == (2x+2y+2)>>2 == avg3-form with taps (x,y,x)? no: x+2y+x is
// 2x+2y. yes.) so one table of three edge indices per output pixel drives them all.
ZW_HD i32 avg3(i32 x, i32 y, i32 z) { return (x + 2 * y + z + 2) >> 2; }

#define ZW_T(a, b, c) (u16)((a) | ((b) << 4) | ((c) << 8))
// index helpers into e[]:  L3=0 L2=1 L1=2 L0=3 P=4 A0=5 .. A7=12
#define ZW_A3(x) ZW_T((x), (x) + 1, (x) + 2)          /* avg3(e[x], e[x+1], e[x+2]) */
#define ZW_A2(x) ZW_T((x), (x) + 1, (x))              /* avg2(e[x], e[x+1])         */
#define ZW_CP(x) ZW_T((x), (x), (x))                  /* copy e[x]                  */
// VE: avg3(p,a0,a1) avg3(a0,a1,a2) avg3(a1,a2,a3) avg3(a2,a3,a4) in every row
// HE: rows avg3(p,l0,l1) avg3(l0,l1,l2) avg3(l1,l2,l3) avg3(l2,l3,l3); with e reversed: p=4,l0=3,l1=2,l2=1,l3=0
//     avg3(p,l0,l1) = taps (4,3,2); avg3(l0,l1,l2) = (3,2,1); avg3(l1,l2,l3) = (2,1,0); avg3(l2,l3,l3) = (1,0,0)
#define ZW_PRED_TABLE_INIT                                                                            \
  { /* mode 2 VE */                                                                                   \
    {ZW_A3(4), ZW_A3(5), ZW_A3(6), ZW_A3(7), ZW_A3(4), ZW_A3(5), ZW_A3(6), ZW_A3(7),                  \
     ZW_A3(4), ZW_A3(5), ZW_A3(6), ZW_A3(7), ZW_A3(4), ZW_A3(5), ZW_A3(6), ZW_A3(7)},                 \
    /* mode 3 HE */                                                                                   \
    {ZW_T(4, 3, 2), ZW_T(4, 3, 2), ZW_T(4, 3, 2), ZW_T(4, 3, 2), ZW_T(3, 2, 1), ZW_T(3, 2, 1),        \
     ZW_T(3, 2, 1), ZW_T(3, 2, 1), ZW_T(2, 1, 0), ZW_T(2, 1, 0), ZW_T(2, 1, 0), ZW_T(2, 1, 0),        \
     ZW_T(1, 0, 0), ZW_T(1, 0, 0), ZW_T(1, 0, 0), ZW_T(1, 0, 0)},                                     \
    /* mode 4 LD: out[y][x] = avgs[y+x], avgs[k] = avg3(a_k, a_k+1, a_k+2), last = avg3(a6,a7,a7) */  \
    {ZW_A3(5), ZW_A3(6), ZW_A3(7), ZW_A3(8), ZW_A3(6), ZW_A3(7), ZW_A3(8), ZW_A3(9),                  \
     ZW_A3(7), ZW_A3(8), ZW_A3(9), ZW_A3(10), ZW_A3(8), ZW_A3(9), ZW_A3(10), ZW_T(11, 12, 12)},       \
    /* mode 5 RD: out[y][x] = avgs[3-y+x], avgs[k] = avg3(e_k, e_k+1, e_k+2) over e0..e8 = idx 0..8 */\
    {ZW_A3(3), ZW_A3(4), ZW_A3(5), ZW_A3(6), ZW_A3(2), ZW_A3(3), ZW_A3(4), ZW_A3(5),                  \
     ZW_A3(1), ZW_A3(2), ZW_A3(3), ZW_A3(4), ZW_A3(0), ZW_A3(1), ZW_A3(2), ZW_A3(3)},                 \
    /* mode 6 VR */                                                                                   \
    {ZW_A2(4), ZW_A2(5), ZW_A2(6), ZW_A2(7), ZW_A3(3), ZW_A3(4), ZW_A3(5), ZW_A3(6),                  \
     ZW_A3(2), ZW_A2(4), ZW_A2(5), ZW_A2(6), ZW_A3(1), ZW_A3(3), ZW_A3(4), ZW_A3(5)},                 \
    /* mode 7 VL */                                                                                   \
    {ZW_A2(5), ZW_A2(6), ZW_A2(7), ZW_A2(8), ZW_A3(5), ZW_A3(6), ZW_A3(7), ZW_A3(8),                  \
     ZW_A2(6), ZW_A2(7), ZW_A2(8), ZW_A3(9), ZW_A3(6), ZW_A3(7), ZW_A3(8), ZW_A3(10)},                \
    /* mode 8 HD */                                                                                   \
    {ZW_A2(3), ZW_A3(3), ZW_A3(4), ZW_A3(5), ZW_A2(2), ZW_A3(2), ZW_A2(3), ZW_A3(3),                  \
     ZW_A2(1), ZW_A3(1), ZW_A2(2), ZW_A3(2), ZW_A2(0), ZW_A3(0), ZW_A2(1), ZW_A3(1)},                 \
    /* mode 9 HU: l0=3 l1=2 l2=1 l3=0 */                                                              \
    {ZW_T(3, 2, 3), ZW_T(3, 2, 1), ZW_T(2, 1, 2), ZW_T(2, 1, 0), ZW_T(2, 1, 2), ZW_T(2, 1, 0),        \
     ZW_T(1, 0, 1), ZW_T(1, 0, 0), ZW_T(1, 0, 1), ZW_T(1, 0, 0), ZW_CP(0), ZW_CP(0),                  \
     ZW_CP(0), ZW_CP(0), ZW_CP(0), ZW_CP(0)},                                                         \
  }

// Single pixel of predictor `mode` (0..9) at raster position k (0..15) from edges e[13].
ZW_HD i32 predict4_pixel(const u8* e, int mode, int k, const u16 (*tab)[16]) {
  if (mode == 0) {
    i32 v = 4 + e[5] + e[6] + e[7] + e[8] + e[0] + e[1] + e[2] + e[3];
    return v >> 3;
  }
  if (mode == 1) {
    i32 l = e[3 - (k >> 2)];
    return clip255(l - (i32)e[4] + (i32)e[5 + (k & 3)]);
  }
  u16 t = tab[mode - 2][k];
  return avg3(e[t & 15], e[(t >> 4) & 15], e[(t >> 8) & 15]);
}

// The same predictors as two-level lookups (used by the cooperative I4 search in zw_search.cuh):
// the pixels of the eight directional modes take only 23 distinct 3-tap values of the edges;
// ZW_DTAPS_INIT[k] = taps (a | b << 4 | c << 8) of value k = (e[a] + 2 e[b] + e[c] + 2) >> 2,
// slot 23 holds the DC prediction, ZW_PRED_IDX_INIT[mode][pixel] selects the slot.
#define ZW_DTAPS_INIT                                                                                          \
  {528, 801, 1074, 1347, 1620, 1893, 2166, 2439, 2712, 2985, 3258, 3275, 256, 16, 289, 562,                    \
   835, 1108, 1381, 1654, 1927, 2200, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define ZW_PRED_IDX_INIT                                                                \
  {                                                                                     \
    {23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23}, /* DC */          \
    {23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23, 23}, /* TM: computed */ \
    {4, 5, 6, 7, 4, 5, 6, 7, 4, 5, 6, 7, 4, 5, 6, 7},                 /* VE */          \
    {2, 2, 2, 2, 1, 1, 1, 1, 0, 0, 0, 0, 12, 12, 12, 12},             /* HE */          \
    {5, 6, 7, 8, 6, 7, 8, 9, 7, 8, 9, 10, 8, 9, 10, 11},              /* LD */          \
    {3, 4, 5, 6, 2, 3, 4, 5, 1, 2, 3, 4, 0, 1, 2, 3},                 /* RD */          \
    {17, 18, 19, 20, 3, 4, 5, 6, 2, 17, 18, 19, 1, 3, 4, 5},          /* VR */          \
    {18, 19, 20, 21, 5, 6, 7, 8, 19, 20, 21, 9, 6, 7, 8, 10},         /* VL */          \
    {16, 3, 4, 5, 15, 2, 16, 3, 14, 1, 15, 2, 13, 0, 14, 1},          /* HD */          \
    {15, 1, 14, 0, 14, 0, 13, 12, 13, 12, 22, 22, 22, 22, 22, 22},    /* HU */          \
  }
// Host-checkable statement of the lookup form (tests/hostcheck compares it with predict4_pixel).
ZW_HD i32 predict4_pixel_lut(const u8* e, int mode, int k, const u16* dtaps, const u8 (*pidx)[16]) {
  if (mode == 1) return clip255((i32)e[3 - (k >> 2)] - (i32)e[4] + (i32)e[5 + (k & 3)]);
  const int slot = pidx[mode][k];
  if (slot == 23) return (4 + e[5] + e[6] + e[7] + e[8] + e[0] + e[1] + e[2] + e[3]) >> 3;
  const u32 t = dtaps[slot];
  return avg3(e[t & 15], e[(t >> 4) & 15], e[(t >> 8) & 15]);
}

}  // namespace zw
#endif  // ZW_PRIMS_CUH
