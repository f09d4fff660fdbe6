// zw_oracle.cpp -- CPU ORACLE: literal C++ restatement of the reference's lossy VP8 encoder.
//
// TEST INFRASTRUCTURE ONLY (see zw_oracle.h).  "parity unpinned" at whole-file level: the
// reference (Rust) cannot be compiled in this image and holds no golden encoder output; the
// component arithmetic is pinned by the reference's own KATs (tests/test_oracle_kats.py).
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference).  The structure deliberately mirrors the Rust code -- including its
// redundant recomputation and its quirks -- so that it can be
// audited side by side.  Compile with -ffp-contract=off (Q15: Rust never fuses mul-add).
#include "zw_oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "vp8_tables.h"

namespace {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i8 = int8_t;
using i16 = int16_t;
using i32 = int32_t;
using i64 = int64_t;

// ---------------------------------------------------------------------------------------------
// src/encoder/fast_math.rs:15-122  (f64 helpers; no FMA contraction)
// ---------------------------------------------------------------------------------------------
static inline u64 f64_bits(double x) { u64 b; memcpy(&b, &x, 8); return b; }
static inline double f64_from_bits(u64 b) { double x; memcpy(&x, &b, 8); return x; }

static double fm_round(double x) { return (double)(i64)(x + 0.5); }  // fast_math.rs:15

static double fm_cbrt(double x) {  // fast_math.rs:22-43
  if (x == 0.0) return 0.0;
  u64 bits = f64_bits(x);
  u64 exp_bias = 1023;
  u64 approx_bits = (bits / 3) + (exp_bias * 2 / 3) * (1ull << 52);
  double y = f64_from_bits(approx_bits);
  for (int i = 0; i < 4; i++) {
    double y2 = y * y;
    y = (2.0 * y + x / y2) / 3.0;
  }
  return y;
}

static double fm_log2(double x) {  // fast_math.rs:67-92
  u64 bits = f64_bits(x);
  i64 exp = (i64)((bits >> 52) & 0x7FF) - 1023;
  u64 mantissa_bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
  double m = f64_from_bits(mantissa_bits);
  double y = (m - 1.0) / (m + 1.0);
  double y2 = y * y;
  const double C0 = 2.8853900817779268;
  const double C1 = 0.9617966939259756;
  const double C2 = 0.5770780163555854;
  const double C3 = 0.4121985831111324;
  const double C4 = 0.3205988987531030;
  double poly = C0 + y2 * (C1 + y2 * (C2 + y2 * (C3 + y2 * C4)));
  double log2_m = y * poly;
  return (double)exp + log2_m;
}

static double fm_exp2(double x) {  // fast_math.rs:96-122
  if (x < -1022.0) x = -1022.0;
  if (x > 1023.0) x = 1023.0;
  i64 xi = (x >= 0.0) ? (i64)x : (i64)x - 1;
  double xf = x - (double)xi;
  const double LN2 = 0.693147180559945309417232121458176568;  // core::f64::consts::LN_2
  const double C0 = 1.0;
  const double C1 = LN2;
  const double C2 = LN2 * LN2 / 2.0;
  const double C3 = LN2 * LN2 * LN2 / 6.0;
  const double C4 = LN2 * LN2 * LN2 * LN2 / 24.0;
  const double C5 = LN2 * LN2 * LN2 * LN2 * LN2 / 120.0;
  double poly = C0 + xf * (C1 + xf * (C2 + xf * (C3 + xf * (C4 + xf * C5))));
  u64 exp_bits = ((u64)(xi + 1023)) << 52;
  double scale = f64_from_bits(exp_bits);
  return poly * scale;
}

static double fm_pow(double x, double n) {  // fast_math.rs:48-62
  if (x <= 0.0) return 0.0;
  if (x == 1.0 || n == 0.0) return 1.0;
  if (n == 1.0) return x;
  return fm_exp2(n * fm_log2(x));
}

// ---------------------------------------------------------------------------------------------
// src/encoder/arithmetic.rs:7-196  (boolean entropy coder)
// ---------------------------------------------------------------------------------------------
struct ArithmeticEncoder {
  std::vector<u8> writer;
  u32 bottom = 0;
  u32 range = 255;
  i32 bit_num = 24;
  std::vector<u16>* log = nullptr;  // dump only: every (bit << 8 | probability) handed to write_bool

  void add_one_to_output() {  // arithmetic.rs:47-60
    size_t i = writer.size();
    while (i > 0) {
      i -= 1;
      if (writer[i] < 255) {
        writer[i] += 1;
        return;
      }
      writer[i] = 0;
    }
    writer.insert(writer.begin(), 1);
  }
  void write_flag(bool f) { write_bool(f, 128); }  // arithmetic.rs:63
  void write_bool(bool b, u8 probability) {        // arithmetic.rs:67-95
    if (log) log->push_back((u16)(((u16)b << 8) | probability));
    u32 split = 1 + (((range - 1) * (u32)probability) >> 8);
    if (b) {
      bottom += split;
      range -= split;
    } else {
      range = split;
    }
    while (range < 128) {
      range <<= 1;
      if (bottom & (1u << 31)) add_one_to_output();
      bottom <<= 1;
      bit_num -= 1;
      if (bit_num == 0) {
        writer.push_back((u8)(bottom >> 24));
        bottom &= (1u << 24) - 1;
        bit_num = 8;
      }
    }
  }
  void write_literal(u8 num_bits, u8 value) {  // arithmetic.rs:97-102
    for (int bit = (int)num_bits - 1; bit >= 0; bit--) {
      bool be = ((1u << bit) & value) > 0;
      write_bool(be, 128);
    }
  }
  void write_optional_signed_value_none() { write_flag(false); }  // arithmetic.rs:104 (only None is used)
  void write_with_tree(const i8* tree, size_t tree_len, const u8* probs, i8 value) {
    write_with_tree_start_index(tree, tree_len, probs, value, 0);
  }
  // arithmetic.rs:120-173
  void write_with_tree_start_index(const i8* tree, size_t tree_len, const u8* probs, i8 value,
                                   size_t start_index) {
    size_t current_index = 0;
    bool found = false;
    for (size_t k = 0; k < tree_len; k++)
      if (tree[k] == (i8)(-value)) { current_index = k; found = true; break; }
    assert(found);
    (void)found;
    bool enc_b[16];
    u8 enc_p[16];
    int count = 0;
    for (;;) {
      if (current_index == start_index) {
        enc_b[count] = false; enc_p[count] = probs[current_index / 2]; count++;
        break;
      }
      if (current_index == start_index + 1) {
        enc_b[count] = true; enc_p[count] = probs[current_index / 2]; count++;
        break;
      }
      bool encode_val;
      if (current_index % 2 == 0) {
        encode_val = false;
      } else {
        current_index -= 1;
        encode_val = true;
      }
      enc_b[count] = encode_val; enc_p[count] = probs[current_index / 2]; count++;
      size_t prev = 0;
      bool ok = false;
      for (size_t k = 0; k < tree_len; k++)
        if (tree[k] == (i8)current_index) { prev = k; ok = true; break; }
      assert(ok);
      (void)ok;
      current_index = prev;
    }
    for (int i = count - 1; i >= 0; i--) write_bool(enc_b[i], enc_p[i]);
  }
  std::vector<u8> flush_and_get_buffer() {  // arithmetic.rs:176-195
    i32 c = bit_num;
    u32 v = bottom;
    if (bottom & (1u << (32 - bit_num))) add_one_to_output();
    v <<= (c & 7);
    c = (c >> 3) - 1;
    while (c >= 0) { v <<= 8; c -= 1; }
    c = 3;
    while (c >= 0) { writer.push_back((u8)(v >> 24)); v <<= 8; c -= 1; }
    return std::move(writer);
  }
};

// src/common/types.rs:191-205,332,686-702 (trees)
const i8 SEGMENT_ID_TREE[6] = {2, 4, 0, -1, -2, -3};
const i8 KEYFRAME_YMODE_TREE[8] = {-4, 2, 4, 6, 0, -1, -2, -3};
const i8 KEYFRAME_BPRED_MODE_TREE[18] = {0, 2, -1, 4, -2, 6, 8, 12, -3, 10, -5, -6, -4, 14, -7, 16, -8, -9};
const i8 KEYFRAME_UV_MODE_TREE[6] = {0, 2, -1, 4, -2, -3};
const i8 DCT_TOKEN_TREE[22] = {-11, 2, 0, 4, -1, 6, 8, 12, -2, 10, -3, -4, 14, 16, -5, -6, 18, 20, -7, -8, -9, -10};
enum { DCT_0 = 0, DCT_1 = 1, DCT_CAT1 = 5, DCT_CAT2 = 6, DCT_CAT3 = 7, DCT_CAT4 = 8, DCT_CAT5 = 9, DCT_CAT6 = 10, DCT_EOB = 11 };
enum { LM_DC = 0, LM_V = 1, LM_H = 2, LM_TM = 3, LM_B = 4 };
enum { IM_DC = 0, IM_TM = 1, IM_VE = 2, IM_HE = 3, IM_LD = 4, IM_RD = 5, IM_VR = 6, IM_VL = 7, IM_HD = 8, IM_HU = 9 };
enum { PLANE_YCOEFF1 = 0, PLANE_Y2 = 1, PLANE_CHROMA = 2, PLANE_YCOEFF0 = 3 };  // types.rs:48-57

typedef u8 TokenProbTables[4][8][3][11];
static inline const u8* coeff_probs_default() { return kCoeffProbs; }

// ---------------------------------------------------------------------------------------------
// src/common/transform.rs
// ---------------------------------------------------------------------------------------------
// ---- primitive-invocation counters (measurement only; SURVEY.md 8(d) "ALGORITHMIC int-ops") ----
// Active only inside the mode-search / transform functions of both passes (OpsGate), i.e. not for
// the re-quantisation the reference repeats for statistics and token emission.
enum { OPC_FDCT, OPC_IDCT, OPC_WHT, OPC_IWHT, OPC_TTRANSFORM, OPC_QUANT_COEFF, OPC_SSE_PX, OPC_COST_COEFF,
       OPC_TRELLIS_POS, OPC_I4_PREDSET, OPC_ADD_RESIDUE, OPC_TRELLIS_BLOCK, OPC_N };
static thread_local int g_ops_gate = 0, g_ops_pass = 0, g_ops_chroma = 0;
static thread_local u64 g_ops[4][OPC_N];  // [pass-1 luma, pass-1 chroma, pass-2 luma, pass-2 chroma]
struct OpsGate { OpsGate() { g_ops_gate++; } ~OpsGate() { g_ops_gate--; } };
struct OpsChroma { int prev; OpsChroma() : prev(g_ops_chroma) { g_ops_chroma = 1; } ~OpsChroma() { g_ops_chroma = prev; } };
#define OPC(i, n) do { if (g_ops_gate) g_ops[2 * g_ops_pass + g_ops_chroma][i] += (u64)(n); } while (0)

static void idct4x4(i32* block) {  // transform.rs:35-79 (scalar; SIMD twin is equivalent, Q16)
  OPC(OPC_IDCT, 1);
  const i64 CONST1 = 20091, CONST2 = 35468;
  for (int i = 0; i < 4; i++) {
    i64 a1 = (i64)block[i] + (i64)block[8 + i];
    i64 b1 = (i64)block[i] - (i64)block[8 + i];
    i64 t1 = ((i64)block[4 + i] * CONST2) >> 16;
    i64 t2 = (i64)block[12 + i] + (((i64)block[12 + i] * CONST1) >> 16);
    i64 c1 = t1 - t2;
    t1 = (i64)block[4 + i] + (((i64)block[4 + i] * CONST1) >> 16);
    t2 = ((i64)block[12 + i] * CONST2) >> 16;
    i64 d1 = t1 + t2;
    block[i] = (i32)(a1 + d1);
    block[4 + i] = (i32)(b1 + c1);
    block[12 + i] = (i32)(a1 - d1);
    block[8 + i] = (i32)(b1 - c1);
  }
  for (int i = 0; i < 4; i++) {
    i64 a1 = (i64)block[4 * i] + (i64)block[4 * i + 2];
    i64 b1 = (i64)block[4 * i] - (i64)block[4 * i + 2];
    i64 t1 = ((i64)block[4 * i + 1] * CONST2) >> 16;
    i64 t2 = (i64)block[4 * i + 3] + (((i64)block[4 * i + 3] * CONST1) >> 16);
    i64 c1 = t1 - t2;
    t1 = (i64)block[4 * i + 1] + (((i64)block[4 * i + 1] * CONST1) >> 16);
    t2 = ((i64)block[4 * i + 3] * CONST2) >> 16;
    i64 d1 = t1 + t2;
    block[4 * i] = (i32)((a1 + d1 + 4) >> 3);
    block[4 * i + 3] = (i32)((a1 - d1 + 4) >> 3);
    block[4 * i + 1] = (i32)((b1 + c1 + 4) >> 3);
    block[4 * i + 2] = (i32)((b1 - c1 + 4) >> 3);
  }
}

static void iwht4x4(i32* block) {  // transform.rs:82-114
  OPC(OPC_IWHT, 1);
  for (int i = 0; i < 4; i++) {
    i32 a1 = block[i] + block[12 + i];
    i32 b1 = block[4 + i] + block[8 + i];
    i32 c1 = block[4 + i] - block[8 + i];
    i32 d1 = block[i] - block[12 + i];
    block[i] = a1 + b1;
    block[4 + i] = c1 + d1;
    block[8 + i] = a1 - b1;
    block[12 + i] = d1 - c1;
  }
  for (int r = 0; r < 4; r++) {
    i32* b = block + 4 * r;
    i32 a1 = b[0] + b[3];
    i32 b1 = b[1] + b[2];
    i32 c1 = b[1] - b[2];
    i32 d1 = b[0] - b[3];
    i32 a2 = a1 + b1, b2 = c1 + d1, c2 = a1 - b1, d2 = d1 - c1;
    b[0] = (a2 + 3) >> 3;
    b[1] = (b2 + 3) >> 3;
    b[2] = (c2 + 3) >> 3;
    b[3] = (d2 + 3) >> 3;
  }
}

static void wht4x4(i32* block) {  // transform.rs:116-158
  OPC(OPC_WHT, 1);
  for (int i = 0; i < 4; i++) {
    i64 a = (i64)block[i * 4] + (i64)block[i * 4 + 3];
    i64 b = (i64)block[i * 4 + 1] + (i64)block[i * 4 + 2];
    i64 c = (i64)block[i * 4 + 1] - (i64)block[i * 4 + 2];
    i64 d = (i64)block[i * 4] - (i64)block[i * 4 + 3];
    block[i * 4] = (i32)(a + b);
    block[i * 4 + 1] = (i32)(c + d);
    block[i * 4 + 2] = (i32)(a - b);
    block[i * 4 + 3] = (i32)(d - c);
  }
  for (int i = 0; i < 4; i++) {
    i64 a1 = (i64)block[i] + (i64)block[i + 12];
    i64 b1 = (i64)block[i + 4] + (i64)block[i + 8];
    i64 c1 = (i64)block[i + 4] - (i64)block[i + 8];
    i64 d1 = (i64)block[i] - (i64)block[i + 12];
    i64 a2 = a1 + b1, b2 = c1 + d1, c2 = a1 - b1, d2 = d1 - c1;
    // Rust `/` truncates toward zero, like C++.
    i64 a3 = (a2 + (a2 > 0 ? 1 : 0)) / 2;
    i64 b3 = (b2 + (b2 > 0 ? 1 : 0)) / 2;
    i64 c3 = (c2 + (c2 > 0 ? 1 : 0)) / 2;
    i64 d3 = (d2 + (d2 > 0 ? 1 : 0)) / 2;
    block[i] = (i32)a3;
    block[i + 4] = (i32)b3;
    block[i + 8] = (i32)c3;
    block[i + 12] = (i32)d3;
  }
}

static void dct4x4(i32* block) {  // transform.rs:176-207 (scalar; SIMD twin is equivalent, Q16)
  OPC(OPC_FDCT, 1);
  for (int i = 0; i < 4; i++) {
    i64 a = ((i64)block[i * 4] + (i64)block[i * 4 + 3]) * 8;
    i64 b = ((i64)block[i * 4 + 1] + (i64)block[i * 4 + 2]) * 8;
    i64 c = ((i64)block[i * 4 + 1] - (i64)block[i * 4 + 2]) * 8;
    i64 d = ((i64)block[i * 4] - (i64)block[i * 4 + 3]) * 8;
    block[i * 4] = (i32)(a + b);
    block[i * 4 + 2] = (i32)(a - b);
    block[i * 4 + 1] = (i32)((c * 2217 + d * 5352 + 14500) >> 12);
    block[i * 4 + 3] = (i32)((d * 2217 - c * 5352 + 7500) >> 12);
  }
  for (int i = 0; i < 4; i++) {
    i64 a = (i64)block[i] + (i64)block[i + 12];
    i64 b = (i64)block[i + 4] + (i64)block[i + 8];
    i64 c = (i64)block[i + 4] - (i64)block[i + 8];
    i64 d = (i64)block[i] - (i64)block[i + 12];
    block[i] = (i32)((a + b + 7) >> 4);
    block[i + 8] = (i32)((a - b + 7) >> 4);
    block[i + 4] = (i32)(((c * 2217 + d * 5352 + 12000) >> 16) + (d != 0 ? 1 : 0));
    block[i + 12] = (i32)((d * 2217 - c * 5352 + 51000) >> 16);
  }
}

// ---------------------------------------------------------------------------------------------
// src/common/prediction.rs
// ---------------------------------------------------------------------------------------------
const size_t LUMA_STRIDE = 32, LUMA_BLOCK_SIZE = 32 * 17;
const size_t CHROMA_STRIDE = 32, CHROMA_BLOCK_SIZE = 32 * 9;
typedef std::array<u8, LUMA_BLOCK_SIZE> LumaBuf;
typedef std::array<u8, CHROMA_BLOCK_SIZE> ChromaBuf;

static LumaBuf create_border_luma(size_t mbx, size_t mby, size_t mbw, const u8* top, const u8* left) {
  // prediction.rs:15-77
  const size_t stride = LUMA_STRIDE;
  LumaBuf ws;
  ws.fill(0);
  {
    u8* above = &ws[1];  // ws[1..stride] : 31 entries
    if (mby == 0) {
      for (size_t i = 0; i < stride - 1; i++) above[i] = 127;
    } else {
      for (size_t i = 0; i < 16; i++) above[i] = top[mbx * 16 + i];
      if (mbx == mbw - 1) {
        for (size_t i = 16; i < stride - 1; i++) above[i] = top[mbx * 16 + 15];
      } else {
        // zip with top[mbx*16+16..] (top holds 16*mbw+4 entries) -> stops at the shorter one
        size_t avail = (16 * mbw + 4) - (mbx * 16 + 16);
        for (size_t i = 16; i < stride - 1 && (i - 16) < avail; i++) above[i] = top[mbx * 16 + i];
      }
    }
  }
  for (size_t i = 17; i < 21; i++) {
    ws[4 * stride + i] = ws[i];
    ws[8 * stride + i] = ws[i];
    ws[12 * stride + i] = ws[i];
  }
  if (mbx == 0) {
    for (size_t i = 0; i < 16; i++) ws[(i + 1) * stride] = 129;
  } else {
    for (size_t i = 0; i < 16; i++) ws[(i + 1) * stride] = left[1 + i];
  }
  ws[0] = (mby == 0) ? 127 : (mbx == 0 ? 129 : left[0]);
  return ws;
}

static ChromaBuf create_border_chroma(size_t mbx, size_t mby, const u8* top, size_t top_len, const u8* left) {
  // prediction.rs:85-126
  const size_t stride = CHROMA_STRIDE;
  ChromaBuf cb;
  cb.fill(0);
  {
    u8* above = &cb[1];
    if (mby == 0) {
      for (size_t i = 0; i < stride - 1; i++) above[i] = 127;
    } else {
      size_t avail = top_len - mbx * 8;  // zip stops at the end of `top`
      for (size_t i = 0; i < stride - 1 && i < avail; i++) above[i] = top[mbx * 8 + i];
    }
  }
  if (mbx == 0) {
    for (size_t y = 0; y < 8; y++) cb[(y + 1) * stride] = 129;
  } else {
    for (size_t y = 0; y < 8; y++) cb[(y + 1) * stride] = left[1 + y];
  }
  cb[0] = (mby == 0) ? 127 : (mbx == 0 ? 129 : left[0]);
  return cb;
}

static void add_residue(u8* pblock, const i32* rblock, size_t y0, size_t x0, size_t stride) {
  OPC(OPC_ADD_RESIDUE, 1);
  // prediction.rs:138-152
  size_t pos = y0 * stride + x0;
  for (int r = 0; r < 4; r++) {
    for (int c = 0; c < 4; c++) {
      i32 v = rblock[r * 4 + c] + (i32)pblock[pos + c];
      pblock[pos + c] = (u8)std::min(std::max(v, 0), 255);
    }
    pos += stride;
  }
}

static inline u8 avg3(u8 l, u8 t, u8 r) { return (u8)(((u16)l + 2 * (u16)t + (u16)r + 2) >> 2); }
static inline u8 avg2(u8 t, u8 r) { return (u8)(((u16)t + (u16)r + 1) >> 1); }

static void predict_vpred(u8* a, size_t size, size_t x0, size_t y0, size_t stride) {  // :164
  for (size_t y = 0; y < size; y++)
    for (size_t x = 0; x < size; x++) a[(y0 + y) * stride + x0 + x] = a[(y0 - 1) * stride + x0 + x];
}
static void predict_hpred(u8* a, size_t size, size_t x0, size_t y0, size_t stride) {  // :175
  for (size_t y = 0; y < size; y++) {
    u8 left = a[(y0 + y) * stride + x0 - 1];
    for (size_t x = 0; x < size; x++) a[(y0 + y) * stride + x0 + x] = left;
  }
}
static void predict_dcpred(u8* a, size_t size, size_t stride, bool above, bool left) {  // :183-212
  u32 sum = 0;
  u32 shf = (size == 8) ? 2 : 3;
  if (left) {
    for (size_t y = 0; y < size; y++) sum += a[(y + 1) * stride];
    shf += 1;
  }
  if (above) {
    for (size_t x = 1; x <= size; x++) sum += a[x];
    shf += 1;
  }
  u32 dcval = (!left && !above) ? 128 : (sum + (1u << (shf - 1))) >> shf;
  for (size_t y = 0; y < size; y++)
    for (size_t x = 0; x < size; x++) a[1 + stride * (y + 1) + x] = (u8)dcval;
}
static void predict_tmpred(u8* a, size_t size, size_t x0, size_t y0, size_t stride) {  // :287-324
  i32 p = a[(y0 - 1) * stride + x0 - 1];
  for (size_t y = 0; y < size; y++) {
    i32 left_minus_p = (i32)a[(y0 + y) * stride + x0 - 1] - p;
    for (size_t x = 0; x < size; x++) {
      i32 v = left_minus_p + (i32)a[(y0 - 1) * stride + x0 + x];
      a[(y0 + y) * stride + x0 + x] = (u8)std::min(std::max(v, 0), 255);
    }
  }
}

// prediction.rs:326-554: the ten 4x4 predictors, written through the edge vector
// e[0..12] = l3 l2 l1 l0 p a0 a1 a2 a3 a4 a5 a6 a7  (same pixels the Rust helpers fetch).
struct Edges { u8 p, a[8], l[4]; };
static Edges fetch_edges(const u8* ws, size_t x0, size_t y0, size_t stride) {
  Edges e;
  e.p = ws[(y0 - 1) * stride + x0 - 1];
  for (int i = 0; i < 8; i++) e.a[i] = ws[(y0 - 1) * stride + x0 + i];
  for (int i = 0; i < 4; i++) e.l[i] = ws[(y0 + i) * stride + x0 - 1];
  return e;
}
static void predict4x4_block(const Edges& E, int mode, u8 out[16]) {
  const u8 p = E.p, a0 = E.a[0], a1 = E.a[1], a2 = E.a[2], a3 = E.a[3], a4 = E.a[4], a5 = E.a[5], a6 = E.a[6], a7 = E.a[7];
  const u8 l0 = E.l[0], l1 = E.l[1], l2 = E.l[2], l3 = E.l[3];
  const u8 e0 = l3, e1 = l2, e2 = l1, e3 = l0, e4 = p, e5 = a0, e6 = a1, e7 = a2, e8 = a3;
  switch (mode) {
    case IM_DC: {  // predict_bdcpred :326-344 / I4Predictions mode 0
      u32 v = 4;
      v += (u32)a0 + a1 + a2 + a3;
      v += (u32)l0 + l1 + l2 + l3;
      u8 dc = (u8)(v >> 3);
      for (int i = 0; i < 16; i++) out[i] = dc;
      break;
    }
    case IM_TM: {  // predict_tmpred(size 4)
      const u8 L[4] = {l0, l1, l2, l3}, A[4] = {a0, a1, a2, a3};
      for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) {
          i32 v = (i32)L[y] - (i32)p + (i32)A[x];
          out[y * 4 + x] = (u8)std::min(std::max(v, 0), 255);
        }
      break;
    }
    case IM_VE: {  // :396-411
      u8 avg[4] = {avg3(p, a0, a1), avg3(a0, a1, a2), avg3(a1, a2, a3), avg3(a2, a3, a4)};
      for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) out[y * 4 + x] = avg[x];
      break;
    }
    case IM_HE: {  // :413-431
      u8 avgs[4] = {avg3(p, l0, l1), avg3(l0, l1, l2), avg3(l1, l2, l3), avg3(l2, l3, l3)};
      for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) out[y * 4 + x] = avgs[y];
      break;
    }
    case IM_LD: {  // :433-452
      u8 avgs[7] = {avg3(a0, a1, a2), avg3(a1, a2, a3), avg3(a2, a3, a4), avg3(a3, a4, a5),
                    avg3(a4, a5, a6), avg3(a5, a6, a7), avg3(a6, a7, a7)};
      for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) out[y * 4 + x] = avgs[y + x];
      break;
    }
    case IM_RD: {  // :454-473
      u8 avgs[7] = {avg3(e0, e1, e2), avg3(e1, e2, e3), avg3(e2, e3, e4), avg3(e3, e4, e5),
                    avg3(e4, e5, e6), avg3(e5, e6, e7), avg3(e6, e7, e8)};
      for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) out[y * 4 + x] = avgs[3 - y + x];
      break;
    }
    case IM_VR: {  // :475-494
      out[12] = avg3(e1, e2, e3); out[8] = avg3(e2, e3, e4);
      out[13] = avg3(e3, e4, e5); out[4] = avg3(e3, e4, e5);
      out[9] = avg2(e4, e5); out[0] = avg2(e4, e5);
      out[14] = avg3(e4, e5, e6); out[5] = avg3(e4, e5, e6);
      out[10] = avg2(e5, e6); out[1] = avg2(e5, e6);
      out[15] = avg3(e5, e6, e7); out[6] = avg3(e5, e6, e7);
      out[11] = avg2(e6, e7); out[2] = avg2(e6, e7);
      out[7] = avg3(e6, e7, e8); out[3] = avg2(e7, e8);
      break;
    }
    case IM_VL: {  // :496-515
      out[0] = avg2(a0, a1); out[4] = avg3(a0, a1, a2);
      out[8] = avg2(a1, a2); out[1] = avg2(a1, a2);
      out[5] = avg3(a1, a2, a3); out[12] = avg3(a1, a2, a3);
      out[9] = avg2(a2, a3); out[2] = avg2(a2, a3);
      out[13] = avg3(a2, a3, a4); out[6] = avg3(a2, a3, a4);
      out[10] = avg2(a3, a4); out[3] = avg2(a3, a4);
      out[14] = avg3(a3, a4, a5); out[7] = avg3(a3, a4, a5);
      out[11] = avg3(a4, a5, a6); out[15] = avg3(a5, a6, a7);
      break;
    }
    case IM_HD: {  // :517-536
      out[12] = avg2(e0, e1); out[13] = avg3(e0, e1, e2);
      out[8] = avg2(e1, e2); out[14] = avg2(e1, e2);
      out[9] = avg3(e1, e2, e3); out[15] = avg3(e1, e2, e3);
      out[10] = avg2(e2, e3); out[4] = avg2(e2, e3);
      out[11] = avg3(e2, e3, e4); out[5] = avg3(e2, e3, e4);
      out[6] = avg2(e3, e4); out[0] = avg2(e3, e4);
      out[7] = avg3(e3, e4, e5); out[1] = avg3(e3, e4, e5);
      out[2] = avg3(e4, e5, e6); out[3] = avg3(e5, e6, e7);
      break;
    }
    default: {  // IM_HU :538-554
      out[0] = avg2(l0, l1); out[1] = avg3(l0, l1, l2);
      out[2] = avg2(l1, l2); out[4] = avg2(l1, l2);
      out[3] = avg3(l1, l2, l3); out[5] = avg3(l1, l2, l3);
      out[6] = avg2(l2, l3); out[8] = avg2(l2, l3);
      out[7] = avg3(l2, l3, l3); out[9] = avg3(l2, l3, l3);
      out[10] = l3; out[11] = l3; out[12] = l3; out[13] = l3; out[14] = l3; out[15] = l3;
      break;
    }
  }
}
// In-place form (vp8.rs:1730-1749 apply_intra4_prediction / transform_luma_blocks_4x4 match arm)
static void apply_intra4_prediction(u8* ws, int mode, size_t x0, size_t y0, size_t stride) {
  Edges e = fetch_edges(ws, x0, y0, stride);
  u8 out[16];
  predict4x4_block(e, mode, out);
  for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) ws[(y0 + y) * stride + x0 + x] = out[y * 4 + x];
}
// I4Predictions::compute (prediction.rs:568-780)
struct I4Predictions { u8 data[10][16]; };
static I4Predictions i4_predictions_compute(const u8* src, size_t x0, size_t y0, size_t stride) {
  OPC(OPC_I4_PREDSET, 1);
  I4Predictions r;
  Edges e = fetch_edges(src, x0, y0, stride);
  for (int m = 0; m < 10; m++) predict4x4_block(e, m, r.data[m]);
  return r;
}

// ---------------------------------------------------------------------------------------------
// src/decoder/yuv.rs:656-899
// ---------------------------------------------------------------------------------------------
const i32 YUV_FIX = 16, YUV_HALF = 1 << 15;
static inline u8 rgb_to_y(const u8* rgb) {  // :859
  i32 luma = 16839 * (i32)rgb[0] + 33059 * (i32)rgb[1] + 6420 * (i32)rgb[2];
  return (u8)((luma + YUV_HALF + (16 << YUV_FIX)) >> YUV_FIX);
}
static inline i32 rgb_to_u_raw(const u8* rgb) {  // :889
  return -9719 * (i32)rgb[0] - 19081 * (i32)rgb[1] + 28800 * (i32)rgb[2] + (128 << YUV_FIX);
}
static inline i32 rgb_to_v_raw(const u8* rgb) {  // :896
  return 28800 * (i32)rgb[0] - 24116 * (i32)rgb[1] - 4684 * (i32)rgb[2] + (128 << YUV_FIX);
}
static inline u8 rgb_to_u_avg(const u8* a, const u8* b, const u8* c, const u8* d) {  // :866
  return (u8)((rgb_to_u_raw(a) + rgb_to_u_raw(b) + rgb_to_u_raw(c) + rgb_to_u_raw(d) + (YUV_HALF << 2)) >> (YUV_FIX + 2));
}
static inline u8 rgb_to_v_avg(const u8* a, const u8* b, const u8* c, const u8* d) {  // :878
  return (u8)((rgb_to_v_raw(a) + rgb_to_v_raw(b) + rgb_to_v_raw(c) + rgb_to_v_raw(d) + (YUV_HALF << 2)) >> (YUV_FIX + 2));
}

static void convert_image_yuv(const u8* image_data, size_t BPP, size_t width, size_t height,
                              std::vector<u8>& y_bytes, std::vector<u8>& u_bytes, std::vector<u8>& v_bytes) {
  // yuv.rs:656-804
  size_t mb_width = (width + 15) / 16, mb_height = (height + 15) / 16;
  size_t luma_width = 16 * mb_width, chroma_width = 8 * mb_width;
  y_bytes.assign(16 * mb_width * 16 * mb_height, 0);
  u_bytes.assign(8 * mb_width * 8 * mb_height, 0);
  v_bytes.assign(8 * mb_width * 8 * mb_height, 0);
  size_t row_pairs = height / 2, col_pairs = width / 2;
  bool odd_height = (height & 1) != 0, odd_width = (width & 1) != 0;
  auto px = [&](size_t r, size_t c) { return image_data + (r * width + c) * BPP; };
  for (size_t row_pair = 0; row_pair < row_pairs; row_pair++) {
    size_t r1 = row_pair * 2, r2 = r1 + 1, cr = row_pair;
    for (size_t col_pair = 0; col_pair < col_pairs; col_pair++) {
      size_t c1 = col_pair * 2, c2 = c1 + 1, cc = col_pair;
      const u8 *p1 = px(r1, c1), *p2 = px(r1, c2), *p3 = px(r2, c1), *p4 = px(r2, c2);
      y_bytes[r1 * luma_width + c1] = rgb_to_y(p1);
      y_bytes[r1 * luma_width + c2] = rgb_to_y(p2);
      y_bytes[r2 * luma_width + c1] = rgb_to_y(p3);
      y_bytes[r2 * luma_width + c2] = rgb_to_y(p4);
      u_bytes[cr * chroma_width + cc] = rgb_to_u_avg(p1, p2, p3, p4);
      v_bytes[cr * chroma_width + cc] = rgb_to_v_avg(p1, p2, p3, p4);
    }
    if (odd_width) {
      size_t c = width - 1, cc = col_pairs;
      const u8 *p1 = px(r1, c), *p3 = px(r2, c);
      y_bytes[r1 * luma_width + c] = rgb_to_y(p1);
      y_bytes[r2 * luma_width + c] = rgb_to_y(p3);
      u_bytes[cr * chroma_width + cc] = rgb_to_u_avg(p1, p1, p3, p3);
      v_bytes[cr * chroma_width + cc] = rgb_to_v_avg(p1, p1, p3, p3);
    }
  }
  if (odd_height) {
    size_t r = height - 1, cr = row_pairs;
    for (size_t col_pair = 0; col_pair < col_pairs; col_pair++) {
      size_t c1 = col_pair * 2, c2 = c1 + 1, cc = col_pair;
      const u8 *p1 = px(r, c1), *p2 = px(r, c2);
      y_bytes[r * luma_width + c1] = rgb_to_y(p1);
      y_bytes[r * luma_width + c2] = rgb_to_y(p2);
      u_bytes[cr * chroma_width + cc] = rgb_to_u_avg(p1, p2, p1, p2);
      v_bytes[cr * chroma_width + cc] = rgb_to_v_avg(p1, p2, p1, p2);
    }
    if (odd_width) {
      size_t c = width - 1, cc = col_pairs;
      const u8* p = px(r, c);
      y_bytes[r * luma_width + c] = rgb_to_y(p);
      u_bytes[cr * chroma_width + cc] = rgb_to_u_avg(p, p, p, p);
      v_bytes[cr * chroma_width + cc] = rgb_to_v_avg(p, p, p, p);
    }
  }
  for (size_t y = 0; y < height; y++) {
    u8 last = y_bytes[y * luma_width + width - 1];
    for (size_t x = width; x < luma_width; x++) y_bytes[y * luma_width + x] = last;
  }
  for (size_t y = height; y < mb_height * 16; y++)
    for (size_t x = 0; x < luma_width; x++) y_bytes[y * luma_width + x] = y_bytes[(height - 1) * luma_width + x];
  size_t chroma_height = (height + 1) / 2, actual_chroma_width = (width + 1) / 2;
  for (size_t y = 0; y < chroma_height; y++) {
    u8 lu = u_bytes[y * chroma_width + actual_chroma_width - 1];
    u8 lv = v_bytes[y * chroma_width + actual_chroma_width - 1];
    for (size_t x = actual_chroma_width; x < chroma_width; x++) {
      u_bytes[y * chroma_width + x] = lu;
      v_bytes[y * chroma_width + x] = lv;
    }
  }
  for (size_t y = chroma_height; y < mb_height * 8; y++)
    for (size_t x = 0; x < chroma_width; x++) {
      u_bytes[y * chroma_width + x] = u_bytes[(chroma_height - 1) * chroma_width + x];
      v_bytes[y * chroma_width + x] = v_bytes[(chroma_height - 1) * chroma_width + x];
    }
}

static void convert_image_y(const u8* image_data, size_t BPP, size_t width, size_t height,
                            std::vector<u8>& y_bytes, std::vector<u8>& u_bytes, std::vector<u8>& v_bytes) {
  // yuv.rs:806-845
  size_t mb_width = (width + 15) / 16, mb_height = (height + 15) / 16;
  size_t luma_width = 16 * mb_width;
  y_bytes.assign(16 * mb_width * 16 * mb_height, 0);
  u_bytes.assign(8 * mb_width * 8 * mb_height, 127);
  v_bytes.assign(8 * mb_width * 8 * mb_height, 127);
  for (size_t y = 0; y < height; y++)
    for (size_t x = 0; x < width; x++) y_bytes[y * luma_width + x] = image_data[(y * width + x) * BPP];
  for (size_t y = 0; y < height; y++) {
    u8 last = y_bytes[y * luma_width + width - 1];
    for (size_t x = width; x < luma_width; x++) y_bytes[y * luma_width + x] = last;
  }
  for (size_t y = height; y < mb_height * 16; y++)
    for (size_t x = 0; x < luma_width; x++) y_bytes[y * luma_width + x] = y_bytes[(height - 1) * luma_width + x];
}

// ---------------------------------------------------------------------------------------------
// src/encoder/cost.rs  (quantiser, costs, trellis, statistics)
// ---------------------------------------------------------------------------------------------
const u32 RD_DISTO_MULT = 256;
const u32 QFIX = 17;
const size_t MAX_LEVEL = 2047, MAX_VARIABLE_LEVEL = 67;
const i32 FLATNESS_LIMIT_I16 = 0, FLATNESS_LIMIT_UV = 2;
const u32 FLATNESS_PENALTY = 140;

static inline u16 vp8_bit_cost(bool bit, u8 prob) {  // cost.rs:40
  return bit ? kEntropyCost[255 - (size_t)prob] : kEntropyCost[(size_t)prob];
}
static inline u32 quantization_bias(u32 b) { return ((b << QFIX) + 128) >> 8; }  // cost.rs:238
static inline i32 quantdiv(u32 coeff, u32 iq, u32 bias) {                         // cost.rs:244
  return (i32)(((u64)coeff * (u64)iq + (u64)bias) >> QFIX);
}

static i32 t_transform(const u8* in, size_t stride, const u16* w) {  // cost.rs:73-107
  OPC(OPC_TTRANSFORM, 1);
  i32 tmp[16];
  for (int i = 0; i < 4; i++) {
    size_t row = i * stride;
    i32 a0 = (i32)in[row] + (i32)in[row + 2];
    i32 a1 = (i32)in[row + 1] + (i32)in[row + 3];
    i32 a2 = (i32)in[row + 1] - (i32)in[row + 3];
    i32 a3 = (i32)in[row] - (i32)in[row + 2];
    tmp[i * 4] = a0 + a1;
    tmp[i * 4 + 1] = a3 + a2;
    tmp[i * 4 + 2] = a3 - a2;
    tmp[i * 4 + 3] = a0 - a1;
  }
  i32 sum = 0;
  for (int i = 0; i < 4; i++) {
    i32 a0 = tmp[i] + tmp[8 + i];
    i32 a1 = tmp[4 + i] + tmp[12 + i];
    i32 a2 = tmp[4 + i] - tmp[12 + i];
    i32 a3 = tmp[i] - tmp[8 + i];
    i32 b0 = a0 + a1, b1 = a3 + a2, b2 = a3 - a2, b3 = a0 - a1;
    sum += (i32)w[i] * std::abs(b0);
    sum += (i32)w[4 + i] * std::abs(b1);
    sum += (i32)w[8 + i] * std::abs(b2);
    sum += (i32)w[12 + i] * std::abs(b3);
  }
  return sum;
}
static i32 tdisto_4x4(const u8* a, const u8* b, size_t stride, const u16* w) {  // cost.rs:122
  i32 s1 = t_transform(a, stride, w), s2 = t_transform(b, stride, w);
  return std::abs(s2 - s1) >> 5;
}
static i32 tdisto_16x16(const u8* a, const u8* b, size_t stride, const u16* w) {  // cost.rs:137
  i32 d = 0;
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) {
      size_t off = y * 4 * stride + x * 4;
      d += tdisto_4x4(a + off, b + off, stride, w);
    }
  return d;
}
static bool is_flat_source_16(const u8* src, size_t stride) {  // cost.rs:177
  u8 v = src[0];
  for (int y = 0; y < 16; y++)
    for (int x = 0; x < 16; x++)
      if (src[y * stride + x] != v) return false;
  return true;
}
static bool is_flat_coeffs(const i16* levels, size_t num_blocks, i32 thresh) {  // cost.rs:199
  i32 score = 0;
  for (size_t b = 0; b < num_blocks; b++)
    for (int i = 1; i < 16; i++)
      if (levels[b * 16 + i] != 0) {
        score += 1;
        if (score > thresh) return false;
      }
  return true;
}

static u8 compute_filter_level(u8 quant_index, u8 sharpness, u8 filter_strength) {  // cost.rs:271-294
  u32 level0 = 5 * (u32)filter_strength;
  u8 qstep = (u8)(kAcTable[quant_index] >> 2);
  size_t pos = std::min<size_t>(qstep, 64 - 1);
  size_t sidx = std::min<size_t>(sharpness, 7);
  u32 base_strength = kLevelsFromDelta[sidx * 64 + pos];
  u32 f = (base_strength * level0) / 256;
  if (f < 2) return 0;
  if (f > 63) return 63;
  return (u8)f;
}

enum MatrixType { MT_Y1 = 0, MT_Y2 = 1, MT_UV = 2 };
struct VP8Matrix {  // cost.rs:386-486
  u16 q[16];
  u32 iq[16];
  u32 bias[16];
  u32 zthresh[16];
  u16 sharpen[16];
  static VP8Matrix make(u16 q_dc, u16 q_ac, MatrixType t) {  // cost.rs:401-447
    u32 b0, b1;
    switch (t) {
      case MT_Y1: b0 = 96; b1 = 110; break;
      case MT_Y2: b0 = 96; b1 = 108; break;
      default: b0 = 110; b1 = 115; break;
    }
    VP8Matrix m;
    memset(&m, 0, sizeof(m));
    m.q[0] = q_dc;
    m.q[1] = q_ac;
    for (int i = 0; i < 2; i++) {
      u32 bias = i > 0 ? b1 : b0;
      m.iq[i] = (u32)((1ull << QFIX) / (u64)m.q[i]);
      m.bias[i] = quantization_bias(bias);
      m.zthresh[i] = ((1u << QFIX) - 1 - m.bias[i]) / m.iq[i];
    }
    for (int i = 2; i < 16; i++) {
      m.q[i] = m.q[1]; m.iq[i] = m.iq[1]; m.bias[i] = m.bias[1]; m.zthresh[i] = m.zthresh[1];
    }
    if (t == MT_Y1)
      for (int i = 0; i < 16; i++) m.sharpen[i] = (u16)(((u32)kFreqSharpening[i] * (u32)m.q[i]) >> 11);
    return m;
  }
  i32 quantize_coeff(i32 coeff, size_t pos) const {  // cost.rs:457
    OPC(OPC_QUANT_COEFF, 1);
    bool sign = coeff < 0;
    u32 abs_coeff = (u32)(sign ? -coeff : coeff);
    i32 level = quantdiv(abs_coeff, iq[pos], bias[pos]);
    return sign ? -level : level;
  }
  i32 dequantize(i32 level, size_t pos) const { return level * (i32)q[pos]; }  // cost.rs:484
};

typedef u16 LevelCostArray[MAX_VARIABLE_LEVEL + 1];
static u16 variable_level_cost(size_t level, const u8* probas) {  // cost.rs:1425-1447
  if (level == 0) return 0;
  size_t idx = std::min(level, MAX_VARIABLE_LEVEL) - 1;
  u16 pattern = kLevelCodes[idx * 2 + 0];
  u16 bits = kLevelCodes[idx * 2 + 1];
  u16 cost = 0;
  u16 p = pattern, b = bits;
  size_t i = 2;
  while (p != 0) {
    if (p & 1) cost += vp8_bit_cost((b & 1) != 0, probas[i]);
    b >>= 1;
    p >>= 1;
    i += 1;
  }
  return cost;
}

struct LevelCosts {  // cost.rs:1452-1591
  u16 level_cost[4][8][3][MAX_VARIABLE_LEVEL + 1];
  size_t remapped[4][16][3];
  u16 eob_cost[4][8][3];
  u16 init_cost[4][8][3];
  bool dirty;
  LevelCosts() {  // ::new  (:1477): everything zero, dirty
    memset(level_cost, 0, sizeof(level_cost));
    memset(remapped, 0, sizeof(remapped));
    memset(eob_cost, 0, sizeof(eob_cost));
    memset(init_cost, 0, sizeof(init_cost));
    dirty = true;
  }
  void calculate(const TokenProbTables probs) {  // :1500-1546
    if (!dirty) return;
    for (int ctype = 0; ctype < 4; ctype++) {
      for (int band = 0; band < 8; band++)
        for (int ctx = 0; ctx < 3; ctx++) {
          const u8* p = probs[ctype][band][ctx];
          u16 cost0 = ctx > 0 ? vp8_bit_cost(true, p[0]) : 0;
          u16 cost_base = vp8_bit_cost(true, p[1]) + cost0;
          level_cost[ctype][band][ctx][0] = vp8_bit_cost(false, p[1]) + cost0;
          for (size_t v = 1; v <= MAX_VARIABLE_LEVEL; v++)
            level_cost[ctype][band][ctx][v] = (u16)(cost_base + variable_level_cost(v, p));
          eob_cost[ctype][band][ctx] = vp8_bit_cost(false, p[0]);
          init_cost[ctype][band][ctx] = vp8_bit_cost(true, p[0]);
        }
      for (int n = 0; n < 16; n++) {
        size_t band = kEncBands[n];
        for (int ctx = 0; ctx < 3; ctx++) remapped[ctype][n][ctx] = band;
      }
    }
    dirty = false;
  }
  u32 get_level_cost(size_t ctype, size_t position, size_t ctx, size_t level) const {  // :1551
    u32 fixed = kLevelFixedCosts[std::min(level, MAX_LEVEL)];
    size_t band = remapped[ctype][position][ctx];
    u32 variable = level_cost[ctype][band][ctx][std::min(level, MAX_VARIABLE_LEVEL)];
    return fixed + variable;
  }
  const u16* get_cost_table(size_t ctype, size_t position, size_t ctx) const {  // :1560
    size_t band = remapped[ctype][position][ctx];
    return level_cost[ctype][band][ctx];
  }
  u16 get_eob_cost(size_t ctype, size_t position, size_t ctx) const {  // :1569
    size_t next_pos = std::min<size_t>(position + 1, 15);
    size_t band = kEncBands[next_pos];
    return eob_cost[ctype][band][ctx];
  }
  u16 get_skip_eob_cost(size_t ctype, size_t first, size_t ctx) const {  // :1579
    return eob_cost[ctype][kEncBands[first]][ctx];
  }
  u16 get_init_cost(size_t ctype, size_t first, size_t ctx) const {  // :1587
    return init_cost[ctype][kEncBands[first]][ctx];
  }
};

// cost.rs:788-1006
static const i64 MAX_COST = INT64_MAX / 2;
static inline i64 rd_score_trellis(u32 lambda, i64 rate, i64 distortion) {
  return rate * (i64)lambda + (i64)RD_DISTO_MULT * distortion;
}
static inline u32 level_cost_fast(const u16* costs, size_t level) {  // :756
  u32 fixed = kLevelFixedCosts[level];
  u32 variable = costs[std::min(level, MAX_VARIABLE_LEVEL)];
  u32 sign_cost = level > 0 ? 256 : 0;
  return fixed + variable + sign_cost;
}
static bool trellis_quantize_block(i32* coeffs, i32* out, const VP8Matrix& mtx, u32 lambda, size_t first,
                                   const LevelCosts& level_costs, size_t ctype, size_t ctx0) {
  struct Node { i8 prev; bool sign; i16 level; };
  struct SS { i64 score; const u16* costs; };
  Node nodes[16][2];
  memset(nodes, 0, sizeof(nodes));
  SS score_states[2][2];
  for (auto& a : score_states) for (auto& s : a) { s.score = MAX_COST; s.costs = nullptr; }
  size_t ss_cur = 0, ss_prev = 1;

  i32 thresh = (i32)((i64)mtx.q[1] * (i64)mtx.q[1] / 4);
  i32 last = (i32)first - 1;
  for (int n = 15; n >= (int)first; n--) {
    size_t j = kZigzag[n];
    i32 err = coeffs[j] * coeffs[j];
    if (err > thresh) { last = n; break; }
  }
  if (last < 15) last += 1;
  OPC(OPC_TRELLIS_BLOCK, 1);
  OPC(OPC_TRELLIS_POS, last - (i32)first + 1);

  i32 best_path[3] = {-1, -1, -1};
  i64 skip_cost = level_costs.get_skip_eob_cost(ctype, first, ctx0);
  i64 best_score = rd_score_trellis(lambda, skip_cost, 0);

  const u16* initial_costs = level_costs.get_cost_table(ctype, first, ctx0);
  i64 init_rate = ctx0 == 0 ? (i64)level_costs.get_init_cost(ctype, first, ctx0) : 0;
  i64 init_score = rd_score_trellis(lambda, init_rate, 0);
  for (int d = 0; d < 2; d++) { score_states[ss_cur][d].score = init_score; score_states[ss_cur][d].costs = initial_costs; }

  for (size_t n = first; (i32)n <= last; n++) {
    size_t j = kZigzag[n];
    i32 q = mtx.q[j];
    u32 iq = mtx.iq[j];
    u32 neutral_bias = quantization_bias(0x00);
    bool sign = coeffs[j] < 0;
    i32 abs_coeff = sign ? -coeffs[j] : coeffs[j];
    i32 coeff_with_sharpen = abs_coeff + (i32)mtx.sharpen[j];
    i32 level0 = std::min(quantdiv((u32)coeff_with_sharpen, iq, neutral_bias), (i32)MAX_LEVEL);
    u32 thresh_bias = quantization_bias(0x80);
    i32 thresh_level = std::min(quantdiv((u32)coeff_with_sharpen, iq, thresh_bias), (i32)MAX_LEVEL);
    std::swap(ss_cur, ss_prev);
    for (int delta = 0; delta < 2; delta++) {
      Node& node = nodes[n][delta];
      i32 level = level0 + delta;
      size_t ctx = std::min<size_t>((size_t)level, 2);
      const u16* next_costs = (n + 1 < 16) ? level_costs.get_cost_table(ctype, n + 1, ctx) : nullptr;
      score_states[ss_cur][delta].score = MAX_COST;
      score_states[ss_cur][delta].costs = next_costs;
      if (level < 0 || level > thresh_level) continue;
      i32 new_error = coeff_with_sharpen - level * q;
      i64 orig_error_sq = (i64)(coeff_with_sharpen * coeff_with_sharpen);
      i64 new_error_sq = (i64)(new_error * new_error);
      i64 weight = kWeightTrellis[j];
      i64 delta_distortion = weight * (new_error_sq - orig_error_sq);
      i64 base_score = rd_score_trellis(lambda, 0, delta_distortion);
      size_t lu = (size_t)level;
      const SS* sp = score_states[ss_prev];
      i64 cost0 = sp[0].costs ? (i64)level_cost_fast(sp[0].costs, lu) : (i64)kLevelFixedCosts[lu];
      i64 score0 = sp[0].score + cost0 * (i64)lambda;
      i64 cost1 = sp[1].costs ? (i64)level_cost_fast(sp[1].costs, lu) : (i64)kLevelFixedCosts[lu];
      i64 score1 = sp[1].score + cost1 * (i64)lambda;
      i64 best_cur_score;
      i8 best_prev;
      if (score1 < score0) { best_cur_score = score1 + base_score; best_prev = 1; }
      else { best_cur_score = score0 + base_score; best_prev = 0; }
      node.sign = sign;
      node.level = (i16)level;
      node.prev = best_prev;
      score_states[ss_cur][delta].score = best_cur_score;
      if (level != 0 && best_cur_score < best_score) {
        i64 eob_cost = n < 15 ? (i64)level_costs.get_eob_cost(ctype, n, ctx) : 0;
        i64 terminal_score = best_cur_score + rd_score_trellis(lambda, eob_cost, 0);
        if (terminal_score < best_score) {
          best_score = terminal_score;
          best_path[0] = (i32)n;
          best_path[1] = delta;
          best_path[2] = best_prev;
        }
      }
    }
  }
  if (first == 1) {
    for (int i = 1; i < 16; i++) { out[i] = 0; coeffs[i] = 0; }
  } else {
    for (int i = 0; i < 16; i++) { out[i] = 0; coeffs[i] = 0; }
  }
  if (best_path[0] == -1) return false;
  bool has_nz = false;
  size_t best_node_delta = (size_t)best_path[1];
  size_t n = (size_t)best_path[0];
  nodes[n][best_node_delta].prev = (i8)best_path[2];
  for (;;) {
    const Node& node = nodes[n][best_node_delta];
    size_t j = kZigzag[n];
    i32 level = node.sign ? -(i32)node.level : (i32)node.level;
    out[n] = level;
    has_nz |= level != 0;
    coeffs[j] = level * (i32)mtx.q[j];
    if (n == first) break;
    best_node_delta = (size_t)node.prev;
    n -= 1;
  }
  return has_nz;
}

static inline u64 rd_score(u32 sse, u16 mode_cost, u32 lambda) {  // cost.rs:1092
  return (u64)sse * (u64)RD_DISTO_MULT + (u64)mode_cost * (u64)lambda;
}
static inline u16 get_i4_mode_cost(size_t top, size_t left, size_t mode) {  // cost.rs:1150
  return kFixedCostsI4[(top * 10 + left) * 10 + mode];
}

struct ProbaStats {  // cost.rs:1173-1255
  u32 stats[4][8][3][11];
  void reset() { memset(stats, 0, sizeof(stats)); }
  void record(size_t t, size_t b, size_t c, size_t p, bool bit) {  // :1200
    u32& s = stats[t][b][c][p];
    if (s >= 0xfffe0000u) s = ((s + 1) >> 1) & 0x7fff7fffu;
    s += 0x00010000u + (bit ? 1u : 0u);
  }
  u8 calc_proba(size_t t, size_t b, size_t c, size_t p) const {  // :1213
    u32 s = stats[t][b][c][p];
    u32 nb = s & 0xffff, total = s >> 16;
    if (total == 0) return 255;
    return (u8)(255 - (nb * 255 / total));
  }
  static i32 branch_cost(i32 nb, i32 total, u8 proba) {  // :1260
    i32 cost_1 = kEntropyCost[255 - (size_t)proba];
    i32 cost_0 = kEntropyCost[(size_t)proba];
    return nb * cost_1 + (total - nb) * cost_0;
  }
  void should_update(size_t t, size_t b, size_t c, size_t p, u8 old_proba, u8 update_proba,
                     bool& upd, u8& new_p, i32& savings) const {  // :1226-1254
    u32 s = stats[t][b][c][p];
    i32 nb = (i32)(s & 0xffff), total = (i32)(s >> 16);
    if (total == 0) { upd = false; new_p = old_proba; savings = 0; return; }
    new_p = calc_proba(t, b, c, p);
    i32 old_cost = branch_cost(nb, total, old_proba) + (i32)vp8_bit_cost(false, update_proba);
    i32 new_cost = branch_cost(nb, total, new_p) + (i32)vp8_bit_cost(true, update_proba) + 8 * 256;
    savings = old_cost - new_cost;
    upd = savings > 0;
  }
};
enum { TT_I16AC = 0, TT_I16DC = 1, TT_CHROMA = 2, TT_I4 = 3 };  // cost.rs:1273

static void record_coeffs(const i32* coeffs, int token_type, size_t first, size_t ctx, ProbaStats& stats) {
  // cost.rs:1297-1397
  size_t t = (size_t)token_type;
  size_t n = first;
  size_t context = ctx;
  i32 last = -1;
  for (int i = 15; i >= 0; i--) if (coeffs[i] != 0) { last = i; break; }
  size_t end_of_block = last >= 0 ? (size_t)(last + 1) : 0;
  if (end_of_block <= first) {
    stats.record(t, kEncBands[first], context, 0, false);
    return;
  }
  bool skip_eob = false;
  while (n < end_of_block) {
    size_t band = kEncBands[n];
    u32 v = (u32)std::abs(coeffs[n]);
    n += 1;
    if (!skip_eob) stats.record(t, band, context, 0, true);
    if (v == 0) {
      stats.record(t, band, context, 1, false);
      skip_eob = true;
      context = 0;
      continue;
    }
    stats.record(t, band, context, 1, true);
    if (v == 1) {
      stats.record(t, band, context, 2, false);
      context = 1;
    } else {
      stats.record(t, band, context, 2, true);
      v = std::min<u32>(v, MAX_VARIABLE_LEVEL);
      if (v <= 4) {
        stats.record(t, band, context, 3, false);
        if (v == 2) {
          stats.record(t, band, context, 4, false);
        } else {
          stats.record(t, band, context, 4, true);
          stats.record(t, band, context, 5, v == 4);
        }
      } else if (v <= 10) {
        stats.record(t, band, context, 3, true);
        stats.record(t, band, context, 6, false);
        stats.record(t, band, context, 7, v > 6);
      } else {
        stats.record(t, band, context, 3, true);
        stats.record(t, band, context, 6, true);
        if (v < 3 + (8 << 2)) {
          stats.record(t, band, context, 8, false);
          stats.record(t, band, context, 9, v >= 3 + (8 << 1));
        } else {
          stats.record(t, band, context, 8, true);
          stats.record(t, band, context, 10, v >= 3 + (8 << 3));
        }
      }
      context = 2;
    }
    // skip_eob intentionally not cleared after a non-zero (Q10)
  }
  if (n < 16) stats.record(t, kEncBands[n], context, 0, false);
}

// cost.rs:1670-1729 (scalar; the SSE2 twin is equivalent for |level| < 32768, Q16)
static u32 get_residual_cost(size_t ctx0, const i32* coeffs, size_t ctype, size_t first,
                             const LevelCosts& costs, const TokenProbTables probs) {
  i32 last = -1;  // Residual::new :1609
  for (int i = 15; i >= 0; i--) if (coeffs[i] != 0) { last = i; break; }
  size_t n = first;
  size_t band = kEncBands[n];
  u8 p0 = probs[ctype][band][ctx0][0];
  size_t ctx = ctx0;
  u32 cost = ctx0 == 0 ? (u32)vp8_bit_cost(true, p0) : 0;
  if (last < 0) return (u32)vp8_bit_cost(false, p0);
  OPC(OPC_COST_COEFF, last - (i32)first + 1);
  while ((i32)n < last) {
    size_t v = (size_t)std::abs(coeffs[n]);
    cost += costs.get_level_cost(ctype, n, ctx, v);
    ctx = v >= 2 ? 2 : v;
    n += 1;
  }
  {
    size_t v = (size_t)std::abs(coeffs[n]);
    cost += costs.get_level_cost(ctype, n, ctx, v);
    if (n < 15) {
      size_t next_band = kEncBands[n + 1];
      size_t next_ctx = v == 1 ? 1 : 2;
      u8 last_p0 = probs[ctype][next_band][next_ctx][0];
      cost += (u32)vp8_bit_cost(false, last_p0);
    }
  }
  return cost;
}
static u32 get_cost_luma4(const i32* levels, bool top_nz, bool left_nz, const LevelCosts& costs,
                          const TokenProbTables probs, bool& has_nz) {  // cost.rs:1906
  size_t ctx = (size_t)top_nz + (size_t)left_nz;
  i32 last = -1;
  for (int i = 15; i >= 0; i--) if (levels[i] != 0) { last = i; break; }
  has_nz = last >= 0;
  return get_residual_cost(ctx, levels, 3, 0, costs, probs);
}
static u32 get_cost_luma16(const i32* dc_levels, const i32 ac_levels[16][16], const LevelCosts& costs,
                           const TokenProbTables probs) {  // cost.rs:1936
  u32 total = 0;
  total += get_residual_cost(0, dc_levels, 1, 0, costs, probs);
  for (int b = 0; b < 16; b++) total += get_residual_cost(0, ac_levels[b], 0, 1, costs, probs);
  return total;
}
static u32 get_cost_uv(const i32 uv_levels[8][16], const LevelCosts& costs, const TokenProbTables probs) {
  u32 total = 0;  // cost.rs:1969
  for (int b = 0; b < 8; b++) total += get_residual_cost(0, uv_levels[b], 2, 0, costs, probs);
  return total;
}

// ---------------------------------------------------------------------------------------------
// src/common/types.rs:761-853  Segment
// ---------------------------------------------------------------------------------------------
struct Segment {
  i16 ydc = 0, yac = 0, y2dc = 0, y2ac = 0, uvdc = 0, uvac = 0;
  i8 quantizer_level = 0;
  u8 quant_index = 0;
  VP8Matrix y1_matrix, y2_matrix, uv_matrix;
  u32 lambda_trellis_i4 = 0, lambda_trellis_i16 = 0, lambda_trellis_uv = 0;
  u32 lambda_i16 = 0, lambda_i4 = 0, lambda_uv = 0, lambda_mode = 0, tlambda = 0;
  void init_matrices() {  // types.rs:806-853
    y1_matrix = VP8Matrix::make((u16)ydc, (u16)yac, MT_Y1);
    y2_matrix = VP8Matrix::make((u16)y2dc, (u16)y2ac, MT_Y2);
    uv_matrix = VP8Matrix::make((u16)uvdc, (u16)uvac, MT_UV);
    u32 q_i4 = ((u32)ydc + 15 * (u32)yac + 8) >> 4;
    u32 q_i16 = ((u32)y2dc + 15 * (u32)y2ac + 8) >> 4;
    u32 q_uv = ((u32)uvdc + 15 * (u32)uvac + 8) >> 4;
    lambda_trellis_i4 = std::max<u32>((7 * q_i4 * q_i4) >> 3, 1);
    lambda_trellis_i16 = std::max<u32>((q_i16 * q_i16) >> 2, 1);
    lambda_trellis_uv = std::max<u32>((q_uv * q_uv) << 1, 1);
    lambda_i4 = std::max<u32>((3 * q_i4 * q_i4) >> 7, 1);
    lambda_i16 = std::max<u32>(3 * q_i16 * q_i16, 1);
    lambda_uv = std::max<u32>((3 * q_uv * q_uv) >> 6, 1);
    lambda_mode = std::max<u32>((q_i4 * q_i4) >> 7, 1);
    tlambda = (50u * q_i4) >> 5;
  }
};
static Segment make_segment(u8 idx, i8 delta) {  // vp8.rs:2336-2347 / 2459-2470
  Segment s;
  s.ydc = kDcQuant[idx];
  s.yac = kAcQuant[idx];
  s.y2dc = (i16)(kDcQuant[idx] * 2);
  s.y2ac = std::max<i16>((i16)((i32)kAcQuant[idx] * 155 / 100), 8);
  s.uvdc = kDcQuant[idx];
  s.uvac = kAcQuant[idx];
  s.quantizer_level = delta;
  s.quant_index = idx;
  s.init_matrices();
  return s;
}

// ---------------------------------------------------------------------------------------------
// src/encoder/analysis.rs
// ---------------------------------------------------------------------------------------------
const size_t BPS = 32;
const size_t Y_OFF_ENC = 0, U_OFF_ENC = 16, V_OFF_ENC = 24;
const size_t VP8_DSP_SCAN[24] = {
    0 + 0 * BPS, 4 + 0 * BPS, 8 + 0 * BPS, 12 + 0 * BPS, 0 + 4 * BPS, 4 + 4 * BPS, 8 + 4 * BPS, 12 + 4 * BPS,
    0 + 8 * BPS, 4 + 8 * BPS, 8 + 8 * BPS, 12 + 8 * BPS, 0 + 12 * BPS, 4 + 12 * BPS, 8 + 12 * BPS, 12 + 12 * BPS,
    0 + 0 * BPS, 4 + 0 * BPS, 0 + 4 * BPS, 4 + 4 * BPS, 8 + 0 * BPS, 12 + 0 * BPS, 8 + 4 * BPS, 12 + 4 * BPS};
const size_t I16DC16 = 0, I16TM16 = 16;
const size_t C8DC8 = 2 * 16 * BPS, C8TM8 = C8DC8 + 16;
const size_t PRED_SIZE_ENC = 32 * BPS + 16 * BPS + 8 * BPS, YUV_SIZE_ENC = BPS * 16;

static void forward_dct_4x4(const u8* src, const u8* pred, size_t ss, size_t ps, i16* out) {  // analysis.rs:172
  i32 tmp[16];
  for (int i = 0; i < 4; i++) {
    i32 d0 = (i32)src[i * ss] - (i32)pred[i * ps];
    i32 d1 = (i32)src[i * ss + 1] - (i32)pred[i * ps + 1];
    i32 d2 = (i32)src[i * ss + 2] - (i32)pred[i * ps + 2];
    i32 d3 = (i32)src[i * ss + 3] - (i32)pred[i * ps + 3];
    i32 a0 = d0 + d3, a1 = d1 + d2, a2 = d1 - d2, a3 = d0 - d3;
    tmp[0 + i * 4] = (a0 + a1) * 8;
    tmp[2 + i * 4] = (a0 - a1) * 8;
    tmp[1 + i * 4] = (a2 * 2217 + a3 * 5352 + 1812) >> 9;
    tmp[3 + i * 4] = (a3 * 2217 - a2 * 5352 + 937) >> 9;
  }
  for (int i = 0; i < 4; i++) {
    i32 a0 = tmp[0 + i] + tmp[12 + i];
    i32 a1 = tmp[4 + i] + tmp[8 + i];
    i32 a2 = tmp[4 + i] - tmp[8 + i];
    i32 a3 = tmp[0 + i] - tmp[12 + i];
    out[0 + i] = (i16)((a0 + a1 + 7) >> 4);
    out[8 + i] = (i16)((a0 - a1 + 7) >> 4);
    out[4 + i] = (i16)(((a2 * 2217 + a3 * 5352 + 12000) >> 16) + (a3 != 0 ? 1 : 0));
    out[12 + i] = (i16)((a3 * 2217 - a2 * 5352 + 51000) >> 16);
  }
}
struct DctHistogram { u32 max_value; size_t last_non_zero; };
static i32 histo_alpha(const DctHistogram& h) {  // analysis.rs:160
  return h.max_value > 1 ? (i32)(510u * (u32)h.last_non_zero / h.max_value) : 0;
}
static DctHistogram collect_histogram_with_offset(const u8* src_buf, size_t src_base, const u8* pred_buf,
                                                  size_t pred_base, size_t start_block, size_t end_block) {
  // analysis.rs:922-947 + from_distribution :139
  u32 distribution[32] = {0};
  for (size_t j = start_block; j < end_block; j++) {
    size_t so = VP8_DSP_SCAN[j];
    i16 dct_out[16];
    forward_dct_4x4(src_buf + src_base + so, pred_buf + pred_base + so, BPS, BPS, dct_out);
    for (int k = 0; k < 16; k++) {
      size_t v = (size_t)((u16)std::abs((int)dct_out[k]) >> 3);
      distribution[std::min<size_t>(v, 31)] += 1;
    }
  }
  DctHistogram h{0, 1};
  for (size_t k = 0; k < 32; k++)
    if (distribution[k] > 0) {
      if (distribution[k] > h.max_value) h.max_value = distribution[k];
      h.last_non_zero = k;
    }
  return h;
}
static void fill_block(u8* dst, u8 v, size_t size) {
  for (size_t y = 0; y < size; y++) for (size_t x = 0; x < size; x++) dst[y * BPS + x] = v;
}
static void an_vertical_pred(u8* dst, const u8* top, size_t size) {  // analysis.rs:303
  if (top) { for (size_t y = 0; y < size; y++) for (size_t x = 0; x < size; x++) dst[y * BPS + x] = top[x]; }
  else fill_block(dst, 127, size);
}
static void an_horizontal_pred(u8* dst, const u8* left, size_t size) {  // analysis.rs:316
  if (left) { for (size_t y = 0; y < size; y++) for (size_t x = 0; x < size; x++) dst[y * BPS + x] = left[y]; }
  else fill_block(dst, 129, size);
}
static void an_pred_dc(u8* dst, const u8* left, const u8* top, size_t size) {  // analysis.rs:259 / 378
  u32 round = (u32)size, shift = size == 16 ? 5 : 4;
  u8 dc_val;
  if (top && left) {
    u32 dc = 0;
    for (size_t i = 0; i < size; i++) dc += (u32)top[i] + (u32)left[i];
    dc_val = (u8)((dc + round) >> shift);
  } else if (top) {
    u32 dc = 0;
    for (size_t i = 0; i < size; i++) dc += top[i];
    dc += dc;
    dc_val = (u8)((dc + round) >> shift);
  } else if (left) {
    u32 dc = 0;
    for (size_t i = 0; i < size; i++) dc += left[i];
    dc += dc;
    dc_val = (u8)((dc + round) >> shift);
  } else {
    dc_val = 0x80;
  }
  fill_block(dst, dc_val, size);
}
static void an_pred_tm(u8* dst, const u8* left_with_corner, const u8* top, size_t size) {  // analysis.rs:332 / 425
  if (left_with_corner && top) {
    i32 tl = left_with_corner[0];
    for (size_t y = 0; y < size; y++) {
      i32 l = left_with_corner[1 + y];
      for (size_t x = 0; x < size; x++) {
        i32 v = l + (i32)top[x] - tl;
        dst[y * BPS + x] = (u8)std::min(std::max(v, 0), 255);
      }
    }
  } else if (left_with_corner) {
    an_horizontal_pred(dst, left_with_corner + 1, size);
  } else if (top) {
    an_vertical_pred(dst, top, size);
  } else {
    fill_block(dst, 129, size);
  }
}
static void import_block(const u8* src, size_t src_stride, u8* dst, size_t w, size_t h, size_t size) {  // :496
  for (size_t y = 0; y < h; y++) {
    for (size_t x = 0; x < w; x++) dst[y * BPS + x] = src[y * src_stride + x];
    if (w < size) {
      u8 lp = dst[y * BPS + w - 1];
      for (size_t x = w; x < size; x++) dst[y * BPS + x] = lp;
    }
  }
  for (size_t y = h; y < size; y++) for (size_t x = 0; x < size; x++) dst[y * BPS + x] = dst[(h - 1) * BPS + x];
}
static void import_line(const u8* src, size_t src_stride, u8* dst, size_t len, size_t total_len) {  // :519
  for (size_t i = 0; i < len; i++) dst[i] = src[i * src_stride];
  u8 last = dst[len > 0 ? len - 1 : 0];
  for (size_t i = len; i < total_len; i++) dst[i] = last;
}

static void analyze_image(const u8* y_src, const u8* u_src, const u8* v_src, size_t width, size_t height,
                          size_t y_stride, size_t uv_stride, std::vector<u8>& mb_alphas, u32 alpha_histogram[256]) {
  // analysis.rs:964-995 with AnalysisIterator (:532-918) inlined
  size_t mb_w = (width + 15) / 16, mb_h = (height + 15) / 16;
  std::vector<u8> yuv_in(YUV_SIZE_ENC, 0), yuv_p(PRED_SIZE_ENC, 0);
  u8 y_left[17], u_left[9], v_left[9];
  std::vector<u8> y_top(mb_w * 16 + 4, 127), uv_top(mb_w * 16, 127);
  mb_alphas.assign(mb_w * mb_h, 0);
  memset(alpha_histogram, 0, 256 * sizeof(u32));
  auto init_left = [&](size_t yy) {
    u8 corner = yy > 0 ? 129 : 127;
    y_left[0] = u_left[0] = v_left[0] = corner;
    for (int i = 1; i < 17; i++) y_left[i] = 129;
    for (int i = 1; i < 9; i++) { u_left[i] = 129; v_left[i] = 129; }
  };
  size_t mb_idx = 0;
  for (size_t y = 0; y < mb_h; y++) {
    init_left(y);
    for (size_t x = 0; x < mb_w; x++) {
      // import (:622-743)
      size_t y_offset = y * 16 * y_stride + x * 16;
      size_t uv_offset = y * 8 * uv_stride + x * 8;
      size_t w = std::min<size_t>(width - x * 16, 16), h = std::min<size_t>(height - y * 16, 16);
      size_t uv_w = (w + 1) / 2, uv_h = (h + 1) / 2;
      import_block(y_src + y_offset, y_stride, &yuv_in[Y_OFF_ENC], w, h, 16);
      import_block(u_src + uv_offset, uv_stride, &yuv_in[U_OFF_ENC], uv_w, uv_h, 8);
      import_block(v_src + uv_offset, uv_stride, &yuv_in[V_OFF_ENC], uv_w, uv_h, 8);
      if (x == 0) {
        init_left(y);
      } else {
        if (y == 0) {
          y_left[0] = u_left[0] = v_left[0] = 127;
        } else {
          y_left[0] = y_src[y_offset - 1 - y_stride];
          u_left[0] = u_src[uv_offset - 1 - uv_stride];
          v_left[0] = v_src[uv_offset - 1 - uv_stride];
        }
        import_line(y_src + y_offset - 1, y_stride, y_left + 1, h, 16);
        import_line(u_src + uv_offset - 1, uv_stride, u_left + 1, uv_h, 8);
        import_line(v_src + uv_offset - 1, uv_stride, v_left + 1, uv_h, 8);
      }
      if (y == 0) {
        for (int i = 0; i < 16; i++) y_top[x * 16 + i] = 127;
        for (int i = 0; i < 8; i++) { uv_top[x * 16 + i] = 127; uv_top[x * 16 + 8 + i] = 127; }
      } else {
        for (size_t i = 0; i < w; i++) y_top[x * 16 + i] = y_src[y_offset - y_stride + i];
        u8 last_y = y_top[x * 16 + w - 1];
        for (size_t i = w; i < 16; i++) y_top[x * 16 + i] = last_y;
        for (size_t i = 0; i < uv_w; i++) {
          uv_top[x * 16 + i] = u_src[uv_offset - uv_stride + i];
          uv_top[x * 16 + 8 + i] = v_src[uv_offset - uv_stride + i];
        }
        u8 last_u = uv_top[x * 16 + uv_w - 1], last_v = uv_top[x * 16 + 8 + uv_w - 1];
        for (size_t i = uv_w; i < 8; i++) { uv_top[x * 16 + i] = last_u; uv_top[x * 16 + 8 + i] = last_v; }
      }
      bool has_left = x > 0, has_top = y > 0;
      // analyze_best_intra16_mode (:811-857)
      i32 best_alpha = -1;
      {
        const u8* lwc = has_left ? y_left : nullptr;
        const u8* top = has_top ? &y_top[x * 16] : nullptr;
        an_pred_dc(&yuv_p[I16DC16], lwc ? lwc + 1 : nullptr, top, 16);
        an_pred_tm(&yuv_p[I16TM16], lwc, top, 16);
        const size_t offs[2] = {I16DC16, I16TM16};
        for (int mode = 0; mode < 2; mode++) {
          DctHistogram hh = collect_histogram_with_offset(yuv_in.data(), Y_OFF_ENC, yuv_p.data(), offs[mode], 0, 16);
          i32 alpha = histo_alpha(hh);
          if (alpha > best_alpha) best_alpha = alpha;
        }
      }
      // analyze_best_uv_mode (:861-917)
      i32 best_uv_alpha = -1;
      {
        const u8* ul = has_left ? u_left : nullptr;
        const u8* vl = has_left ? v_left : nullptr;
        const u8* ut = has_top ? &uv_top[x * 16] : nullptr;
        const u8* vt = has_top ? &uv_top[x * 16 + 8] : nullptr;
        an_pred_dc(&yuv_p[C8DC8], ul ? ul + 1 : nullptr, ut, 8);
        an_pred_dc(&yuv_p[C8DC8 + 8], vl ? vl + 1 : nullptr, vt, 8);
        an_pred_tm(&yuv_p[C8TM8], ul, ut, 8);
        an_pred_tm(&yuv_p[C8TM8 + 8], vl, vt, 8);
        const size_t offs[2] = {C8DC8, C8TM8};
        for (int mode = 0; mode < 2; mode++) {
          DctHistogram hh = collect_histogram_with_offset(yuv_in.data(), U_OFF_ENC, yuv_p.data(), offs[mode], 16, 24);
          i32 alpha = histo_alpha(hh);
          if (alpha > best_uv_alpha) best_uv_alpha = alpha;
        }
      }
      i32 alpha = (3 * best_alpha + best_uv_alpha + 2) >> 2;  // analyze_macroblock :951
      alpha = 255 - alpha;                                     // final_alpha_value :248
      u8 a8 = (u8)std::min(std::max(alpha, 0), 255);
      mb_alphas[mb_idx] = a8;
      alpha_histogram[a8] += 1;
      mb_idx += 1;
    }
  }
}

static void assign_segments_kmeans(const u32 alphas[256], size_t num_segments, u8 centers[4], u8 map[256],
                                   i32& weighted_average_out) {  // analysis.rs:1029-1130
  num_segments = std::min<size_t>(num_segments, 4);
  memset(centers, 0, 4);
  memset(map, 0, 256);
  size_t min_a = 0, max_a = 255;
  for (size_t n = 0; n < 256; n++) if (alphas[n] > 0) { min_a = n; break; }
  for (size_t n = 255;; n--) {
    if (alphas[n] > 0) { max_a = n; break; }
    if (n == min_a) break;
  }
  size_t range_a = max_a >= min_a ? max_a - min_a : 0;
  for (size_t k = 0; k < num_segments; k++) {
    size_t n = 1 + 2 * k;
    centers[k] = (u8)(min_a + (n * range_a) / (2 * num_segments));
  }
  u32 accum[4] = {0}, dist_accum[4] = {0};
  i32 weighted_average = 0;
  u32 total_weight = 0;
  for (int iter = 0; iter < 6; iter++) {
    for (size_t i = 0; i < num_segments; i++) { accum[i] = 0; dist_accum[i] = 0; }
    size_t current_center = 0;
    for (size_t a = min_a; a <= max_a; a++) {
      if (alphas[a] > 0) {
        while (current_center + 1 < num_segments) {
          i32 d_curr = std::abs((i32)a - (i32)centers[current_center]);
          i32 d_next = std::abs((i32)a - (i32)centers[current_center + 1]);
          if (d_next < d_curr) current_center += 1; else break;
        }
        map[a] = (u8)current_center;
        dist_accum[current_center] += (u32)a * alphas[a];
        accum[current_center] += alphas[a];
      }
    }
    i32 displaced = 0;
    weighted_average = 0;
    total_weight = 0;
    for (size_t n = 0; n < num_segments; n++) {
      if (accum[n] > 0) {
        u8 new_center = (u8)((dist_accum[n] + accum[n] / 2) / accum[n]);
        displaced += std::abs((i32)centers[n] - (i32)new_center);
        centers[n] = new_center;
        weighted_average += (i32)new_center * (i32)accum[n];
        total_weight += accum[n];
      }
    }
    if (displaced < 5) break;
  }
  if (total_weight > 0) weighted_average = (weighted_average + (i32)total_weight / 2) / (i32)total_weight;
  else weighted_average = 128;
  for (size_t i = num_segments; i < 4; i++) centers[i] = centers[num_segments - 1];
  weighted_average_out = weighted_average;
}

static u8 compute_segment_quant(u8 base_quant, i32 segment_alpha, u8 sns_strength) {  // analysis.rs:1145-1174
  const double SNS_TO_DQ = 0.9;
  double amp = SNS_TO_DQ * (double)sns_strength / 100.0 / 128.0;
  double expn = 1.0 - amp * (double)segment_alpha;
  if (expn <= 0.0) return base_quant;
  double c_base = 1.0 - ((double)base_quant / 127.0);
  double c = fm_pow(c_base, expn);
  i32 q = (i32)(127.0 * (1.0 - c));
  return (u8)std::min(std::max(q, 0), 127);
}

// vp8.rs:37-55
static u8 quality_to_quant_index(u8 quality) {
  double c = (double)quality / 100.0;
  double linear_c = c < 0.75 ? c * (2.0 / 3.0) : 2.0 * c - 1.0;
  double cc = fm_cbrt(linear_c);
  i32 q = (i32)fm_round(127.0 * (1.0 - cc));
  return (u8)std::min(std::max(q, 0), 127);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// stage dumps
// ---------------------------------------------------------------------------------------------
struct zwo_dump {
  std::map<std::string, std::vector<u8>> stages;
  template <class T>
  void put(const std::string& name, const T* p, size_t n) {
    const u8* b = reinterpret_cast<const u8*>(p);
    stages[name].assign(b, b + n * sizeof(T));
  }
};

namespace {

// ---------------------------------------------------------------------------------------------
// src/encoder/vp8.rs  Vp8Encoder
// ---------------------------------------------------------------------------------------------
struct Complexity {  // vp8.rs:130-146
  u8 y2 = 0, y[4] = {0, 0, 0, 0}, u[2] = {0, 0}, v[2] = {0, 0};
  void clear(bool include_y2) {
    memset(y, 0, 4); memset(u, 0, 2); memset(v, 0, 2);
    if (include_y2) y2 = 0;
  }
  u16 pack() const {
    u16 r = y2 ? 1 : 0;
    for (int i = 0; i < 4; i++) r |= (u16)(y[i] ? 1 : 0) << (1 + i);
    for (int i = 0; i < 2; i++) r |= (u16)(u[i] ? 1 : 0) << (5 + i);
    for (int i = 0; i < 2; i++) r |= (u16)(v[i] ? 1 : 0) << (7 + i);
    return r;
  }
};
struct MacroblockInfo {  // vp8.rs:160-172
  int luma_mode = LM_DC;
  bool has_bpred = false;
  int luma_bpred[16] = {0};
  int chroma_mode = LM_DC;
  int segment_id = -1;  // None
  bool coeffs_skipped = false;
};

struct Vp8Encoder {
  std::vector<u8>* writer;
  // Frame (types.rs:12-45)
  u16 width = 0, height = 0;
  std::vector<u8> ybuf, ubuf, vbuf;
  u8 filter_level = 0, sharpness_level = 0;
  bool filter_type = false;

  ArithmeticEncoder encoder;
  Segment segments[4];
  bool segments_enabled = false, segments_update_map = false;
  u8 segment_tree_probs[3] = {255, 255, 255};
  std::vector<u8> segment_map;
  bool loop_filter_adjustments = false;
  bool has_no_skip = false;
  u8 macroblock_no_skip_coeff = 0;
  u8 yac_abs = 0;
  TokenProbTables token_probs;
  ProbaStats proba_stats;
  bool has_updated_probs = false;
  TokenProbTables updated_probs;
  LevelCosts level_costs;
  bool do_trellis = true, do_error_diffusion = true;
  u8 method = 4;
  std::vector<Complexity> top_complexity;
  Complexity left_complexity;
  std::vector<int> top_b_pred;
  int left_b_pred[4] = {0, 0, 0, 0};
  u16 macroblock_width = 0, macroblock_height = 0;
  std::vector<ArithmeticEncoder> partitions;
  u8 left_border_y[17], left_border_u[9], left_border_v[9];
  std::vector<u8> top_border_y, top_border_u, top_border_v;
  std::vector<std::array<std::array<i8, 2>, 2>> top_derr;
  i8 left_derr[2][2] = {{0, 0}, {0, 0}};

  zwo_dump* dump = nullptr;
  std::vector<zwo_mb_record> rec_p1, rec_p2;

  explicit Vp8Encoder(std::vector<u8>* w) : writer(w) {
    partitions.emplace_back();
    memset(token_probs, 0, sizeof(token_probs));
    memset(updated_probs, 0, sizeof(updated_probs));
    proba_stats.reset();
    memset(left_border_y, 0, sizeof(left_border_y));
    memset(left_border_u, 0, sizeof(left_border_u));
    memset(left_border_v, 0, sizeof(left_border_v));
  }

  const Segment& get_segment_for_mb(size_t mbx, size_t mby) const {  // vp8.rs:293
    size_t id = 0;
    if (segments_enabled && !segment_map.empty()) id = segment_map[mby * macroblock_width + mbx];
    return segments[id];
  }
  int get_segment_id_for_mb(size_t mbx, size_t mby) const {  // vp8.rs:305
    if (segments_enabled && !segment_map.empty()) return segment_map[mby * macroblock_width + mbx];
    return -1;
  }
  const u8 (*cur_probs() const)[8][3][11] {  // `updated_probs.as_ref().unwrap_or(&token_probs)`
    return has_updated_probs ? updated_probs : token_probs;
  }

  void write_u24_le(u32 v) { writer->push_back((u8)v); writer->push_back((u8)(v >> 8)); writer->push_back((u8)(v >> 16)); }
  void write_u16_le(u16 v) { writer->push_back((u8)v); writer->push_back((u8)(v >> 8)); }

  void write_uncompressed_frame_header(u32 partition_size) {  // vp8.rs:315-330
    u32 version = 0, for_display = 1, keyframe_bit = 0;
    u32 tag = (partition_size << 5) | (for_display << 4) | (version << 1) | keyframe_bit;
    write_u24_le(tag);
    writer->push_back(0x9d); writer->push_back(0x01); writer->push_back(0x2a);
    write_u16_le(width & 0x3FFF);
    write_u16_le(height & 0x3FFF);
  }

  void encode_segment_updates() {  // vp8.rs:393-437
    encoder.write_flag(segments_update_map);
    bool update_data = segments_enabled;
    encoder.write_flag(update_data);
    if (update_data) {
      encoder.write_flag(false);
      for (auto& seg : segments) {
        bool has_delta = seg.quantizer_level != 0;
        encoder.write_flag(has_delta);
        if (has_delta) {
          u8 abs_val = (u8)std::abs((int)seg.quantizer_level);
          encoder.write_literal(7, abs_val);
          encoder.write_flag(seg.quantizer_level < 0);
        }
      }
      for (int i = 0; i < 4; i++) encoder.write_flag(false);
    }
    if (segments_update_map)
      for (u8 prob : segment_tree_probs) {
        bool has_prob = prob != 255;
        encoder.write_flag(has_prob);
        if (has_prob) encoder.write_literal(8, prob);
      }
  }
  void encode_updated_token_probabilities() {  // vp8.rs:462-496
    bool have = has_updated_probs;
    TokenProbTables upd;
    memcpy(upd, updated_probs, sizeof(upd));
    has_updated_probs = false;  // .take()
    for (int t = 0; t < 4; t++)
      for (int b = 0; b < 8; b++)
        for (int c = 0; c < 3; c++)
          for (int p = 0; p < 11; p++) {
            u8 update_prob = kCoeffUpdateProbs[((t * 8 + b) * 3 + c) * 11 + p];
            u8 old_prob = token_probs[t][b][c][p];
            bool should_update = false;
            u8 new_prob = old_prob;
            if (have) { new_prob = upd[t][b][c][p]; should_update = new_prob != old_prob; }
            if (should_update) {
              encoder.write_bool(true, update_prob);
              encoder.write_literal(8, new_prob);
              token_probs[t][b][c][p] = new_prob;
            } else {
              encoder.write_bool(false, update_prob);
            }
          }
  }
  void encode_compressed_frame_header() {  // vp8.rs:332-372
    encoder.write_literal(1, 0);
    encoder.write_literal(1, 0);
    encoder.write_flag(segments_enabled);
    if (segments_enabled) encode_segment_updates();
    encoder.write_flag(filter_type);
    encoder.write_literal(6, filter_level);
    encoder.write_literal(3, sharpness_level);
    encoder.write_flag(loop_filter_adjustments);
    if (loop_filter_adjustments) encoder.write_flag(false);
    u8 partitions_value = 0;  // ilog2(1)
    encoder.write_literal(2, partitions_value);
    // encode_quantization_indices (:445-458)
    encoder.write_literal(7, yac_abs);
    for (int i = 0; i < 5; i++) encoder.write_optional_signed_value_none();
    encoder.write_literal(1, 0);  // refresh entropy probs
    encode_updated_token_probabilities();
    encoder.write_literal(1, has_no_skip ? 1 : 0);
    if (has_no_skip) encoder.write_literal(8, macroblock_no_skip_coeff);
  }
  void write_partitions() {  // vp8.rs:374-391 (always exactly one token partition, D2)
    std::vector<ArithmeticEncoder> parts = std::move(partitions);
    partitions.clear();
    std::vector<std::vector<u8>> bytes;
    for (auto& p : parts) bytes.push_back(p.flush_and_get_buffer());
    if (bytes.size() > 1)
      for (size_t i = 0; i + 1 < bytes.size(); i++) {
        write_u24_le((u32)bytes[i].size());
        writer->insert(writer->end(), bytes[i].begin(), bytes[i].end());
      }
    if (dump) dump->put("PART1", bytes.back().data(), bytes.back().size());
    writer->insert(writer->end(), bytes.back().begin(), bytes.back().end());
  }

  static int luma_into_intra(int lm) {  // types.rs:114-122
    switch (lm) { case LM_DC: return IM_DC; case LM_V: return IM_VE; case LM_H: return IM_HE; default: return IM_TM; }
  }
  void write_macroblock_header(const MacroblockInfo& info, size_t mbx) {  // vp8.rs:498-560
    if (segments_enabled && segments_update_map) {
      i8 segment_id = (i8)(info.segment_id < 0 ? 0 : info.segment_id);
      encoder.write_with_tree(SEGMENT_ID_TREE, 6, segment_tree_probs, segment_id);
    }
    if (has_no_skip) encoder.write_bool(info.coeffs_skipped, macroblock_no_skip_coeff);
    encoder.write_with_tree(KEYFRAME_YMODE_TREE, 8, kKfYmodeProbs, (i8)info.luma_mode);
    if (info.luma_mode == LM_B) {
      assert(info.has_bpred);
      for (size_t y = 0; y < 4; y++) {
        int left = left_b_pred[y];
        for (size_t x = 0; x < 4; x++) {
          int top = top_b_pred[mbx * 4 + x];
          const u8* probs = &kKfBmodeProbs[((size_t)top * 10 + (size_t)left) * 9];
          int intra_mode = info.luma_bpred[y * 4 + x];
          encoder.write_with_tree(KEYFRAME_BPRED_MODE_TREE, 18, probs, (i8)intra_mode);
          left = intra_mode;
          top_b_pred[mbx * 4 + x] = intra_mode;
        }
        left_b_pred[y] = left;
      }
    } else {
      int im = luma_into_intra(info.luma_mode);
      for (int i = 0; i < 4; i++) { left_b_pred[i] = im; top_b_pred[4 * mbx + i] = im; }
    }
    encoder.write_with_tree(KEYFRAME_UV_MODE_TREE, 6, kKfUvModeProbs, (i8)info.chroma_mode);
  }

  void apply_chroma_error_diffusion(i32* u_blocks, i32* v_blocks, size_t mbx, const VP8Matrix& uv_matrix) {
    // vp8.rs:572-647
    const i32 C1 = 7, C2 = 8, DSHIFT = 4, DSCALE = 1;
    i32 q = (i32)uv_matrix.q[0];
    u32 iq = uv_matrix.iq[0];
    u32 bias = uv_matrix.bias[0];
    auto diffuse_dc = [&](i32& dc, i8 top_err, i8 left_err) -> i8 {
      i32 adjustment = (C1 * (i32)top_err + C2 * (i32)left_err) >> (DSHIFT - DSCALE);
      dc += adjustment;
      bool sign = dc < 0;
      u32 abs_dc = (u32)(sign ? -dc : dc);
      u32 zthresh = ((1u << 17) - 1 - bias) / iq;
      i32 level = abs_dc > zthresh ? (i32)((abs_dc * iq + bias) >> 17) : 0;  // u32 arithmetic (abs_dc*iq < 2^32)
      i32 err = (i32)abs_dc - level * q;
      i32 signed_err = sign ? -err : err;
      i32 r = signed_err >> DSCALE;
      return (i8)std::min(std::max(r, -127), 127);
    };
    auto process_channel = [&](i32* blocks, const i8 top[2], const i8 left[2], i8 errs[4]) {
      errs[0] = diffuse_dc(blocks[0], top[0], left[0]);
      errs[1] = diffuse_dc(blocks[16], top[1], errs[0]);
      errs[2] = diffuse_dc(blocks[32], errs[0], left[1]);
      errs[3] = diffuse_dc(blocks[48], errs[1], errs[2]);
    };
    i8 u_errs[4], v_errs[4];
    i8 t0[2] = {top_derr[mbx][0][0], top_derr[mbx][0][1]}, t1[2] = {top_derr[mbx][1][0], top_derr[mbx][1][1]};
    process_channel(u_blocks, t0, left_derr[0], u_errs);
    process_channel(v_blocks, t1, left_derr[1], v_errs);
    const i8* all[2] = {u_errs, v_errs};
    for (int ch = 0; ch < 2; ch++) {
      i8 err1 = all[ch][1], err2 = all[ch][2], err3 = all[ch][3];
      left_derr[ch][0] = err1;
      left_derr[ch][1] = (i8)((3 * (i32)err3) >> 2);
      top_derr[mbx][ch][0] = err2;
      top_derr[mbx][ch][1] = (i8)((i32)err3 - (i32)left_derr[ch][1]);
    }
  }

  // vp8.rs:798-958.  `rec_levels` receives the zig-zag levels actually coded.
  bool encode_coefficients(const i32* block, size_t partition_index, int plane, size_t complexity,
                           const VP8Matrix& matrix, bool has_trellis, u32 trellis_lambda, i16* rec_levels) {
    ArithmeticEncoder& enc = partitions[partition_index];
    size_t first_coeff = plane == PLANE_YCOEFF1 ? 1 : 0;
    const u8(*probs)[3][11] = token_probs[plane];
    assert(complexity <= 2);
    i32 zigzag_block[16] = {0};
    if (has_trellis) {
      i32 coeffs[16];
      memcpy(coeffs, block, sizeof(coeffs));
      trellis_quantize_block(coeffs, zigzag_block, matrix, trellis_lambda, first_coeff, level_costs, (size_t)plane, complexity);
    } else {
      for (size_t i = first_coeff; i < 16; i++) {
        size_t zi = kZigzag[i];
        zigzag_block[i] = matrix.quantize_coeff(block[zi], zi);
      }
    }
    if (rec_levels) for (int i = 0; i < 16; i++) rec_levels[i] = (i16)zigzag_block[i];
    size_t end_of_block_index = 0;
    for (int i = 15; i >= 0; i--) if (zigzag_block[i] != 0) { end_of_block_index = (size_t)i + 1; break; }
    bool skip_eob = false;
    for (size_t index = first_coeff; index < end_of_block_index; index++) {
      i32 coeff = zigzag_block[index];
      size_t band = kCoeffBands[index];
      const u8* probabilities = probs[band][complexity];
      size_t start_index_token_tree = skip_eob ? 2 : 0;
      i32 a = std::abs(coeff);
      int token;
      if (a == 0) {
        enc.write_with_tree_start_index(DCT_TOKEN_TREE, 22, probabilities, DCT_0, start_index_token_tree);
        skip_eob = true;
        token = DCT_0;
      } else if (a <= 4) {
        enc.write_with_tree_start_index(DCT_TOKEN_TREE, 22, probabilities, (i8)a, start_index_token_tree);
        skip_eob = false;
        token = a;
      } else {
        int category;
        if (a <= 6) category = DCT_CAT1;
        else if (a <= 10) category = DCT_CAT2;
        else if (a <= 18) category = DCT_CAT3;
        else if (a <= 34) category = DCT_CAT4;
        else if (a <= 66) category = DCT_CAT5;
        else { assert(a <= 2048); category = DCT_CAT6; }
        enc.write_with_tree_start_index(DCT_TOKEN_TREE, 22, probabilities, (i8)category, start_index_token_tree);
        const u8* category_probs = &kProbDctCat[(category - DCT_CAT1) * 12];
        i32 extra = a - (i32)kDctCatBase[category - DCT_CAT1];
        i32 mask = category == DCT_CAT6 ? (1 << (11 - 1)) : (1 << (category - DCT_CAT1));
        for (int k = 0; k < 12; k++) {
          u8 prob = category_probs[k];
          if (prob == 0) break;
          bool extra_bool = (extra & mask) > 0;
          enc.write_bool(extra_bool, prob);
          mask >>= 1;
        }
        skip_eob = false;
        token = category;
      }
      if (token != DCT_0) enc.write_flag(!(coeff > 0));
      complexity = token == DCT_0 ? 0 : (token == DCT_1 ? 1 : 2);
    }
    if (end_of_block_index < 16) {
      size_t band_index = std::max(first_coeff, end_of_block_index);
      size_t band = kCoeffBands[band_index];
      enc.write_with_tree(DCT_TOKEN_TREE, 22, probs[band][complexity], DCT_EOB);
    }
    return end_of_block_index > 0;
  }

  static void get_coeffs0_from_block(const i32* blocks, i32* coeffs0) {  // vp8.rs:3124
    for (int i = 0; i < 16; i++) coeffs0[i] = blocks[i * 16];
  }

  void encode_residual_data(const MacroblockInfo& info, size_t partition_index, size_t mbx, const i32* y_block_data,
                            const i32* u_block_data, const i32* v_block_data, zwo_mb_record* rec) {
    // vp8.rs:650-794
    int plane = info.luma_mode == LM_B ? PLANE_YCOEFF0 : PLANE_Y2;
    const Segment& segment = segments[info.segment_id < 0 ? 0 : info.segment_id];
    VP8Matrix y1_matrix = segment.y1_matrix, y2_matrix = segment.y2_matrix, uv_matrix = segment.uv_matrix;
    bool is_i4 = info.luma_mode == LM_B;
    bool y1_trellis = do_trellis;
    u32 y1_trellis_lambda = is_i4 ? segment.lambda_trellis_i4 : segment.lambda_trellis_i16;
    if (plane == PLANE_Y2) {
      i32 coeffs0[16];
      get_coeffs0_from_block(y_block_data, coeffs0);
      wht4x4(coeffs0);
      u8 complexity = left_complexity.y2 + top_complexity[mbx].y2;
      bool has_coeffs = encode_coefficients(coeffs0, partition_index, plane, complexity, y2_matrix, false, 0,
                                            rec ? rec->levels[0] : nullptr);
      left_complexity.y2 = has_coeffs ? 1 : 0;
      top_complexity[mbx].y2 = has_coeffs ? 1 : 0;
      plane = PLANE_YCOEFF1;
    }
    for (size_t y = 0; y < 4; y++) {
      u8 left = left_complexity.y[y];
      for (size_t x = 0; x < 4; x++) {
        const i32* block = y_block_data + y * 4 * 16 + x * 16;
        u8 top = top_complexity[mbx].y[x];
        u8 complexity = left + top;
        bool has_coeffs = encode_coefficients(block, partition_index, plane, complexity, y1_matrix, y1_trellis,
                                              y1_trellis_lambda, rec ? rec->levels[1 + y * 4 + x] : nullptr);
        left = has_coeffs ? 1 : 0;
        top_complexity[mbx].y[x] = has_coeffs ? 1 : 0;
      }
      left_complexity.y[y] = left;
    }
    plane = PLANE_CHROMA;
    for (size_t y = 0; y < 2; y++) {
      u8 left = left_complexity.u[y];
      for (size_t x = 0; x < 2; x++) {
        const i32* block = u_block_data + y * 2 * 16 + x * 16;
        u8 top = top_complexity[mbx].u[x];
        u8 complexity = left + top;
        bool has_coeffs = encode_coefficients(block, partition_index, plane, complexity, uv_matrix, false, 0,
                                              rec ? rec->levels[17 + y * 2 + x] : nullptr);
        left = has_coeffs ? 1 : 0;
        top_complexity[mbx].u[x] = has_coeffs ? 1 : 0;
      }
      left_complexity.u[y] = left;
    }
    for (size_t y = 0; y < 2; y++) {
      u8 left = left_complexity.v[y];
      for (size_t x = 0; x < 2; x++) {
        const i32* block = v_block_data + y * 2 * 16 + x * 16;
        u8 top = top_complexity[mbx].v[x];
        u8 complexity = left + top;
        bool has_coeffs = encode_coefficients(block, partition_index, plane, complexity, uv_matrix, false, 0,
                                              rec ? rec->levels[21 + y * 2 + x] : nullptr);
        left = has_coeffs ? 1 : 0;
        top_complexity[mbx].v[x] = has_coeffs ? 1 : 0;
      }
      left_complexity.v[y] = left;
    }
  }

  bool check_all_coeffs_zero(const MacroblockInfo& info, const i32* y_block_data, const i32* u_block_data,
                             const i32* v_block_data) const {  // vp8.rs:962-1023
    const Segment& segment = segments[info.segment_id < 0 ? 0 : info.segment_id];
    const VP8Matrix &y1 = segment.y1_matrix, &y2 = segment.y2_matrix, &uv = segment.uv_matrix;
    if (info.luma_mode != LM_B) {
      i32 coeffs0[16];
      get_coeffs0_from_block(y_block_data, coeffs0);
      wht4x4(coeffs0);
      for (size_t idx = 0; idx < 16; idx++) if (y2.quantize_coeff(coeffs0[idx], idx) != 0) return false;
      for (size_t b = 0; b < 16; b++)
        for (size_t idx = 1; idx < 16; idx++) if (y1.quantize_coeff(y_block_data[b * 16 + idx], idx) != 0) return false;
    } else {
      for (size_t b = 0; b < 16; b++)
        for (size_t idx = 0; idx < 16; idx++) if (y1.quantize_coeff(y_block_data[b * 16 + idx], idx) != 0) return false;
    }
    for (size_t b = 0; b < 4; b++)
      for (size_t idx = 0; idx < 16; idx++) if (uv.quantize_coeff(u_block_data[b * 16 + idx], idx) != 0) return false;
    for (size_t b = 0; b < 4; b++)
      for (size_t idx = 0; idx < 16; idx++) if (uv.quantize_coeff(v_block_data[b * 16 + idx], idx) != 0) return false;
    return true;
  }

  void record_residual_stats(const MacroblockInfo& info, size_t mbx, const i32* y_block_data, const i32* u_block_data,
                             const i32* v_block_data, zwo_mb_record* rec) {  // vp8.rs:1027-1198
    const Segment& segment = segments[info.segment_id < 0 ? 0 : info.segment_id];
    VP8Matrix y1_matrix = segment.y1_matrix, y2_matrix = segment.y2_matrix, uv_matrix = segment.uv_matrix;
    bool is_i4 = info.luma_mode == LM_B;
    auto save = [&](int slot, const i32* zz) { if (rec) for (int i = 0; i < 16; i++) rec->levels[slot][i] = (i16)zz[i]; };
    if (!is_i4) {
      i32 coeffs0[16];
      get_coeffs0_from_block(y_block_data, coeffs0);
      wht4x4(coeffs0);
      i32 zigzag[16] = {0};
      for (size_t i = 0; i < 16; i++) { size_t zi = kZigzag[i]; zigzag[i] = y2_matrix.quantize_coeff(coeffs0[zi], zi); }
      u8 complexity = left_complexity.y2 + top_complexity[mbx].y2;
      record_coeffs(zigzag, TT_I16DC, 0, std::min<u8>(complexity, 2), proba_stats);
      save(0, zigzag);
      bool has_coeffs = false;
      for (int i = 0; i < 16; i++) has_coeffs |= zigzag[i] != 0;
      left_complexity.y2 = has_coeffs ? 1 : 0;
      top_complexity[mbx].y2 = has_coeffs ? 1 : 0;
    }
    int token_type = is_i4 ? TT_I4 : TT_I16AC;
    size_t first_coeff = is_i4 ? 0 : 1;
    u32 trellis_lambda = is_i4 ? segment.lambda_trellis_i4 : segment.lambda_trellis_i16;
    for (size_t y = 0; y < 4; y++) {
      u8 left = left_complexity.y[y];
      for (size_t x = 0; x < 4; x++) {
        const i32* block = y_block_data + y * 4 * 16 + x * 16;
        i32 zigzag[16] = {0};
        u8 top = top_complexity[mbx].y[x];
        size_t ctx0 = std::min<u8>(left + top, 2);
        if (do_trellis) {
          i32 coeffs[16];
          memcpy(coeffs, block, sizeof(coeffs));
          trellis_quantize_block(coeffs, zigzag, y1_matrix, trellis_lambda, first_coeff, level_costs, (size_t)token_type, ctx0);
        } else {
          for (size_t i = first_coeff; i < 16; i++) { size_t zi = kZigzag[i]; zigzag[i] = y1_matrix.quantize_coeff(block[zi], zi); }
        }
        record_coeffs(zigzag, token_type, first_coeff, ctx0, proba_stats);
        save(1 + (int)(y * 4 + x), zigzag);
        bool has_coeffs = false;
        for (size_t i = first_coeff; i < 16; i++) has_coeffs |= zigzag[i] != 0;
        left = has_coeffs ? 1 : 0;
        top_complexity[mbx].y[x] = has_coeffs ? 1 : 0;
      }
      left_complexity.y[y] = left;
    }
    for (int ch = 0; ch < 2; ch++) {
      const i32* data = ch == 0 ? u_block_data : v_block_data;
      for (size_t y = 0; y < 2; y++) {
        u8 left = ch == 0 ? left_complexity.u[y] : left_complexity.v[y];
        for (size_t x = 0; x < 2; x++) {
          const i32* block = data + y * 2 * 16 + x * 16;
          i32 zigzag[16] = {0};
          for (size_t i = 0; i < 16; i++) { size_t zi = kZigzag[i]; zigzag[i] = uv_matrix.quantize_coeff(block[zi], zi); }
          u8 top = ch == 0 ? top_complexity[mbx].u[x] : top_complexity[mbx].v[x];
          u8 complexity = std::min<u8>(left + top, 2);
          record_coeffs(zigzag, TT_CHROMA, 0, complexity, proba_stats);
          save(17 + ch * 4 + (int)(y * 2 + x), zigzag);
          bool has_coeffs = false;
          for (int i = 0; i < 16; i++) has_coeffs |= zigzag[i] != 0;
          left = has_coeffs ? 1 : 0;
          if (ch == 0) top_complexity[mbx].u[x] = has_coeffs ? 1 : 0; else top_complexity[mbx].v[x] = has_coeffs ? 1 : 0;
        }
        if (ch == 0) left_complexity.u[y] = left; else left_complexity.v[y] = left;
      }
    }
  }

  void compute_updated_probabilities() {  // vp8.rs:1202-1238
    TokenProbTables updated;
    memcpy(updated, token_probs, sizeof(updated));
    i32 total_savings = 0;
    u32 num_updates = 0;
    for (size_t t = 0; t < 4; t++)
      for (size_t b = 0; b < 8; b++)
        for (size_t c = 0; c < 3; c++)
          for (size_t p = 0; p < 11; p++) {
            u8 old_prob = token_probs[t][b][c][p];
            u8 update_prob = kCoeffUpdateProbs[((t * 8 + b) * 3 + c) * 11 + p];
            bool should_update; u8 new_p; i32 savings;
            proba_stats.should_update(t, b, c, p, old_prob, update_prob, should_update, new_p, savings);
            if (should_update && savings > 0) {
              updated[t][b][c][p] = new_p;
              total_savings = (i32)((u32)total_savings + (u32)savings);  // release-mode Rust wraps
              num_updates += 1;
            }
          }
    if (total_savings > 0 && num_updates > 0) {
      memcpy(updated_probs, updated, sizeof(updated));
      has_updated_probs = true;
    } else {
      has_updated_probs = false;
    }
  }

  void reset_for_second_pass() {  // vp8.rs:1241-1279  (top_derr deliberately NOT reset, Q4)
    for (auto& c : top_complexity) c = Complexity();
    left_complexity = Complexity();
    for (auto& p : top_b_pred) p = IM_DC;
    for (int i = 0; i < 4; i++) left_b_pred[i] = IM_DC;
    memset(left_border_y, 129, 17); memset(left_border_u, 129, 9); memset(left_border_v, 129, 9);
    std::fill(top_border_y.begin(), top_border_y.end(), 127);
    std::fill(top_border_u.begin(), top_border_u.end(), 127);
    std::fill(top_border_v.begin(), top_border_v.end(), 127);
    partitions.clear();
    partitions.emplace_back();
    encoder = ArithmeticEncoder();
  }

  // ---- prediction helpers -----------------------------------------------------------------
  LumaBuf get_predicted_luma_block_16x16(int luma_mode, size_t mbx, size_t mby) const {  // vp8.rs:2509-2537
    LumaBuf b = create_border_luma(mbx, mby, macroblock_width, top_border_y.data(), left_border_y);
    switch (luma_mode) {
      case LM_V: predict_vpred(b.data(), 16, 1, 1, LUMA_STRIDE); break;
      case LM_H: predict_hpred(b.data(), 16, 1, 1, LUMA_STRIDE); break;
      case LM_TM: predict_tmpred(b.data(), 16, 1, 1, LUMA_STRIDE); break;
      case LM_DC: predict_dcpred(b.data(), 16, LUMA_STRIDE, mby != 0, mbx != 0); break;
      default: assert(false);
    }
    return b;
  }
  void get_luma_blocks_from_predicted_16x16(const LumaBuf& pred, size_t mbx, size_t mby, i32* luma_blocks) const {
    // vp8.rs:2601-2637 (scalar twin; SIMD ftransform2 is equivalent)
    size_t stride = LUMA_STRIDE, w = (size_t)macroblock_width * 16;
    for (size_t by = 0; by < 4; by++)
      for (size_t bx = 0; bx < 4; bx++) {
        size_t block_index = by * 16 * 4 + bx * 16;
        size_t border_block_index = (by * 4 + 1) * stride + bx * 4 + 1;
        size_t y_data_block_index = (mby * 16 + by * 4) * w + mbx * 16 + bx * 4;
        i32 block[16];
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++)
            block[y * 4 + x] = (i32)ybuf[y_data_block_index + y * w + x] - (i32)pred[border_block_index + y * stride + x];
        dct4x4(block);
        memcpy(luma_blocks + block_index, block, sizeof(block));
      }
  }
  ChromaBuf get_predicted_chroma_block(int chroma_mode, size_t mbx, size_t mby, const std::vector<u8>& top_border,
                                       const u8* left_border) const {  // vp8.rs:2918-2950
    ChromaBuf b = create_border_chroma(mbx, mby, top_border.data(), top_border.size(), left_border);
    switch (chroma_mode) {
      case LM_DC: predict_dcpred(b.data(), 8, CHROMA_STRIDE, mby != 0, mbx != 0); break;
      case LM_V: predict_vpred(b.data(), 8, 1, 1, CHROMA_STRIDE); break;
      case LM_H: predict_hpred(b.data(), 8, 1, 1, CHROMA_STRIDE); break;
      default: predict_tmpred(b.data(), 8, 1, 1, CHROMA_STRIDE); break;
    }
    return b;
  }
  void get_chroma_blocks_from_predicted(const ChromaBuf& pred, const std::vector<u8>& chroma_data, size_t mbx, size_t mby,
                                        i32* chroma_blocks) const {  // vp8.rs:2952-2991
    size_t stride = CHROMA_STRIDE, cw = (size_t)macroblock_width * 8;
    for (size_t by = 0; by < 2; by++)
      for (size_t bx = 0; bx < 2; bx++) {
        size_t block_index = by * 16 * 2 + bx * 16;
        size_t border_block_index = (by * 4 + 1) * stride + bx * 4 + 1;
        size_t data_index = (mby * 8 + by * 4) * cw + mbx * 8 + bx * 4;
        i32 block[16];
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++)
            block[y * 4 + x] = (i32)chroma_data[data_index + y * cw + x] - (i32)pred[border_block_index + y * stride + x];
        dct4x4(block);
        memcpy(chroma_blocks + block_index, block, sizeof(block));
      }
  }

  static u32 sse_16x16_luma(const u8* src_y, size_t src_width, size_t mbx, size_t mby, const LumaBuf& pred) {  // vp8.rs:66
    OPC(OPC_SSE_PX, 256);
    u32 sse = 0;
    size_t src_base = mby * 16 * src_width + mbx * 16;
    for (size_t y = 0; y < 16; y++)
      for (size_t x = 0; x < 16; x++) {
        i32 d = (i32)src_y[src_base + y * src_width + x] - (i32)pred[(y + 1) * LUMA_STRIDE + 1 + x];
        sse += (u32)(d * d);
      }
    return sse;
  }
  static u32 sse_8x8_chroma(const u8* src_uv, size_t src_width, size_t mbx, size_t mby, const ChromaBuf& pred) {  // vp8.rs:97
    OPC(OPC_SSE_PX, 64);
    u32 sse = 0;
    size_t src_base = mby * 8 * src_width + mbx * 8;
    for (size_t y = 0; y < 8; y++)
      for (size_t x = 0; x < 8; x++) {
        i32 d = (i32)src_uv[src_base + y * src_width + x] - (i32)pred[(y + 1) * CHROMA_STRIDE + 1 + x];
        sse += (u32)(d * d);
      }
    return sse;
  }

  // ---- mode search ------------------------------------------------------------------------
  int pick_best_intra16(size_t mbx, size_t mby, u64& score_out) const {  // vp8.rs:1504-1681
    size_t mbw = macroblock_width, src_width = mbw * 16;
    const int MODES[4] = {LM_DC, LM_V, LM_H, LM_TM};
    const Segment& segment = get_segment_for_mb(mbx, mby);
    const VP8Matrix &y1_matrix = segment.y1_matrix, &y2_matrix = segment.y2_matrix;
    u32 lambda = segment.lambda_i16, tlambda = segment.tlambda;
    const u8(*probs)[8][3][11] = cur_probs();
    size_t src_base = mby * 16 * src_width + mbx * 16;
    bool is_flat = is_flat_source_16(&ybuf[src_base], src_width);
    int best_mode = LM_DC;
    i64 best_rd_score = INT64_MAX;
    u32 best_coeff_cost = 0;
    u16 best_mode_cost = 0;
    u32 best_sse = 0;
    i32 best_spectral_disto = 0;
    for (size_t mode_idx = 0; mode_idx < 4; mode_idx++) {
      int mode = MODES[mode_idx];
      if (mode == LM_V && mby == 0) continue;
      if (mode == LM_H && mbx == 0) continue;
      if (mode == LM_TM && (mbx == 0 || mby == 0)) continue;
      LumaBuf pred = get_predicted_luma_block_16x16(mode, mbx, mby);
      i32 luma_blocks[256];
      get_luma_blocks_from_predicted_16x16(pred, mbx, mby, luma_blocks);
      i32 y2_coeffs[16];
      for (int i = 0; i < 16; i++) y2_coeffs[i] = luma_blocks[i * 16];
      wht4x4(y2_coeffs);
      i32 y2_quant[16];
      for (size_t idx = 0; idx < 16; idx++) y2_quant[idx] = y2_matrix.quantize_coeff(y2_coeffs[idx], idx);
      i32 y1_quant[16][16];
      memset(y1_quant, 0, sizeof(y1_quant));
      for (size_t b = 0; b < 16; b++)
        for (size_t i = 1; i < 16; i++) y1_quant[b][i] = y1_matrix.quantize_coeff(luma_blocks[b * 16 + i], i);
      u32 coeff_cost = get_cost_luma16(y2_quant, y1_quant, level_costs, probs);
      i32 y2_dequant[16];
      for (size_t idx = 0; idx < 16; idx++) y2_dequant[idx] = y2_matrix.dequantize(y2_quant[idx], idx);
      iwht4x4(y2_dequant);
      LumaBuf reconstructed = pred;
      for (size_t b = 0; b < 16; b++) {
        size_t bx = b % 4, by = b / 4;
        i32 block[16] = {0};
        for (size_t i = 1; i < 16; i++) block[i] = y1_matrix.dequantize(y1_quant[b][i], i);
        block[0] = y2_dequant[b];
        idct4x4(block);
        add_residue(reconstructed.data(), block, 1 + by * 4, 1 + bx * 4, LUMA_STRIDE);
      }
      u32 sse = sse_16x16_luma(ybuf.data(), src_width, mbx, mby, reconstructed);
      i32 spectral_disto = 0;
      if (tlambda > 0) {
        u8 src_block[256], rec_block[256];
        for (size_t y = 0; y < 16; y++)
          for (size_t x = 0; x < 16; x++) {
            src_block[y * 16 + x] = ybuf[(mby * 16 + y) * src_width + mbx * 16 + x];
            rec_block[y * 16 + x] = reconstructed[(y + 1) * LUMA_STRIDE + x + 1];
          }
        i32 td = tdisto_16x16(src_block, rec_block, 16, kWeightY);
        spectral_disto = ((i32)tlambda * td + 128) >> 8;
      }
      u32 d_final = sse;
      i32 sd_final = spectral_disto;
      if (is_flat) {
        i16 all_levels[256] = {0};
        for (size_t b = 0; b < 16; b++) for (size_t i = 1; i < 16; i++) all_levels[b * 16 + i] = (i16)y1_quant[b][i];
        if (is_flat_coeffs(all_levels, 16, FLATNESS_LIMIT_I16)) { d_final = sse * 2; sd_final = spectral_disto * 2; }
      }
      u16 mode_cost = kFixedCostsI16[mode_idx];
      i64 rate = ((i64)mode_cost + (i64)coeff_cost) * (i64)lambda;
      i64 distortion = (i64)RD_DISTO_MULT * ((i64)d_final + (i64)sd_final);
      i64 rds = rate + distortion;
      if (rds < best_rd_score) {
        best_rd_score = rds; best_mode = mode; best_coeff_cost = coeff_cost; best_mode_cost = mode_cost;
        best_sse = d_final; best_spectral_disto = sd_final;
      }
    }
    u32 lambda_mode = segment.lambda_mode;
    i64 final_rate = ((i64)best_mode_cost + (i64)best_coeff_cost) * (i64)lambda_mode;
    i64 final_distortion = (i64)RD_DISTO_MULT * ((i64)best_sse + (i64)best_spectral_disto);
    i64 final_score = final_rate + final_distortion;
    score_out = (u64)std::max<i64>(final_score, 0);
    return best_mode;
  }

  bool pick_best_intra4(size_t mbx, size_t mby, u64 i16_score, int best_modes[16]) const {  // vp8.rs:1790-2036
    size_t mbw = macroblock_width, src_width = mbw * 16;
    size_t best_mode_indices[16] = {0};
    for (int i = 0; i < 16; i++) best_modes[i] = IM_DC;
    LumaBuf y_with_border = create_border_luma(mbx, mby, mbw, top_border_y.data(), left_border_y);
    const Segment& segment = get_segment_for_mb(mbx, mby);
    u32 lambda_i4 = segment.lambda_i4, lambda_mode = segment.lambda_mode;
    u64 running_score = 211ull * (u64)lambda_mode;
    u32 total_mode_cost = 0;
    const u32 max_header_bits = 256 * 16 * 16 / 4;
    bool top_nz[4] = {false, false, false, false}, left_nz[4] = {false, false, false, false};
    const u8(*probs)[8][3][11] = cur_probs();
    for (size_t sby = 0; sby < 4; sby++)
      for (size_t sbx = 0; sbx < 4; sbx++) {
        size_t i = sby * 4 + sbx, y0 = sby * 4 + 1, x0 = sbx * 4 + 1;
        size_t top_ctx = sby == 0 ? 0 : best_mode_indices[(sby - 1) * 4 + sbx];
        size_t left_ctx = sbx == 0 ? 0 : best_mode_indices[sby * 4 + (sbx - 1)];
        bool nz_top = sby == 0 ? false : top_nz[sbx];
        bool nz_left = sbx == 0 ? false : left_nz[sby];
        int best_mode = IM_DC;
        size_t best_mode_idx = 0;
        u64 best_block_score = UINT64_MAX;
        bool best_has_nz = false;
        i32 best_quantized[16] = {0};
        u32 best_sse = 0, best_rate = 0;
        I4Predictions preds = i4_predictions_compute(y_with_border.data(), x0, y0, LUMA_STRIDE);
        size_t src_base = (mby * 16 + sby * 4) * src_width + mbx * 16 + sbx * 4;
        u8 src_block[16];
        for (size_t y = 0; y < 4; y++) for (size_t x = 0; x < 4; x++) src_block[y * 4 + x] = ybuf[src_base + y * src_width + x];
        const VP8Matrix& y1_matrix = segment.y1_matrix;
        size_t max_modes_to_try = method <= 3 ? 3 : (method == 4 ? 4 : 10);
        std::pair<u32, size_t> mode_sse[10];
        OPC(OPC_SSE_PX, 160);
        for (size_t m = 0; m < 10; m++) {
          u32 sse = 0;
          for (int k = 0; k < 16; k++) { i32 d = (i32)src_block[k] - (i32)preds.data[m][k]; sse += (u32)(d * d); }
          mode_sse[m] = {sse, m};
        }
        // sort_unstable_by_key on 10 elements == insertion sort == stable (Q11)
        std::stable_sort(mode_sse, mode_sse + 10, [](const std::pair<u32, size_t>& a, const std::pair<u32, size_t>& b) { return a.first < b.first; });
        for (size_t k = 0; k < max_modes_to_try; k++) {
          size_t mode_idx = mode_sse[k].second;
          const u8* pred = preds.data[mode_idx];
          i32 residual[16];
          for (int t = 0; t < 16; t++) residual[t] = (i32)src_block[t] - (i32)pred[t];
          dct4x4(residual);
          i32 quantized[16];
          for (size_t idx = 0; idx < 16; idx++) quantized[idx] = y1_matrix.quantize_coeff(residual[idx], idx);
          bool has_nz;
          u32 coeff_cost = get_cost_luma4(quantized, nz_top, nz_left, level_costs, probs, has_nz);
          i32 dequantized[16];
          for (size_t idx = 0; idx < 16; idx++) dequantized[idx] = y1_matrix.dequantize(quantized[idx], idx);
          idct4x4(dequantized);
          OPC(OPC_SSE_PX, 16);
          OPC(OPC_ADD_RESIDUE, 1);
          u32 sse = 0;
          for (int t = 0; t < 16; t++) {
            i32 rec = std::min(std::max((i32)pred[t] + dequantized[t], 0), 255);
            i32 d = (i32)src_block[t] - rec;
            sse += (u32)(d * d);
          }
          u16 mode_cost = get_i4_mode_cost(top_ctx, left_ctx, mode_idx);
          u32 total_rate = (u32)mode_cost + coeff_cost;
          u64 rds = rd_score(sse, (u16)total_rate, lambda_i4);  // u16 truncation, Q8
          if (rds < best_block_score) {
            best_block_score = rds; best_mode = (int)mode_idx; best_mode_idx = mode_idx; best_has_nz = has_nz;
            memcpy(best_quantized, quantized, sizeof(quantized)); best_sse = sse; best_rate = total_rate;
          }
        }
        best_modes[i] = best_mode;
        best_mode_indices[i] = best_mode_idx;
        top_nz[sbx] = best_has_nz;
        left_nz[sby] = best_has_nz;
        u16 mode_cost = get_i4_mode_cost(top_ctx, left_ctx, best_mode_idx);
        total_mode_cost += (u32)mode_cost;
        u64 block_score_for_comparison = rd_score(best_sse, (u16)best_rate, lambda_mode);
        running_score += block_score_for_comparison;
        if (running_score >= i16_score) return false;
        if (total_mode_cost > max_header_bits) return false;
        apply_intra4_prediction(y_with_border.data(), best_mode, x0, y0, LUMA_STRIDE);
        i32 dequantized[16];
        for (size_t idx = 0; idx < 16; idx++) dequantized[idx] = y1_matrix.dequantize(best_quantized[idx], idx);
        idct4x4(dequantized);
        add_residue(y_with_border.data(), dequantized, y0, x0, LUMA_STRIDE);
      }
    return true;
  }

  int pick_best_uv(size_t mbx, size_t mby) const {  // vp8.rs:2050-2200
    OpsChroma ops_chroma_;
    size_t mbw = macroblock_width, chroma_width = mbw * 8;
    const int MODES[4] = {LM_DC, LM_V, LM_H, LM_TM};
    const Segment& segment = get_segment_for_mb(mbx, mby);
    const VP8Matrix& uv_matrix = segment.uv_matrix;
    u32 lambda = segment.lambda_uv;
    const u8(*probs)[8][3][11] = cur_probs();
    int best_mode = LM_DC;
    i64 best_rd_score = INT64_MAX;
    for (size_t mode_idx = 0; mode_idx < 4; mode_idx++) {
      int mode = MODES[mode_idx];
      if (mode == LM_V && mby == 0) continue;
      if (mode == LM_H && mbx == 0) continue;
      if (mode == LM_TM && (mbx == 0 || mby == 0)) continue;
      ChromaBuf pred_u = get_predicted_chroma_block(mode, mbx, mby, top_border_u, left_border_u);
      ChromaBuf pred_v = get_predicted_chroma_block(mode, mbx, mby, top_border_v, left_border_v);
      i32 u_blocks[64], v_blocks[64];
      get_chroma_blocks_from_predicted(pred_u, ubuf, mbx, mby, u_blocks);
      get_chroma_blocks_from_predicted(pred_v, vbuf, mbx, mby, v_blocks);
      i32 uv_quant[8][16];
      for (size_t b = 0; b < 4; b++) for (size_t i = 0; i < 16; i++) uv_quant[b][i] = uv_matrix.quantize_coeff(u_blocks[b * 16 + i], i);
      for (size_t b = 0; b < 4; b++) for (size_t i = 0; i < 16; i++) uv_quant[4 + b][i] = uv_matrix.quantize_coeff(v_blocks[b * 16 + i], i);
      u32 coeff_cost = get_cost_uv(uv_quant, level_costs, probs);
      ChromaBuf rec_u = pred_u, rec_v = pred_v;
      for (size_t b = 0; b < 4; b++) {
        size_t bx = b % 2, by = b / 2;
        i32 block[16];
        for (size_t i = 0; i < 16; i++) block[i] = uv_matrix.dequantize(uv_quant[b][i], i);
        idct4x4(block);
        add_residue(rec_u.data(), block, 1 + by * 4, 1 + bx * 4, CHROMA_STRIDE);
      }
      for (size_t b = 0; b < 4; b++) {
        size_t bx = b % 2, by = b / 2;
        i32 block[16];
        for (size_t i = 0; i < 16; i++) block[i] = uv_matrix.dequantize(uv_quant[4 + b][i], i);
        idct4x4(block);
        add_residue(rec_v.data(), block, 1 + by * 4, 1 + bx * 4, CHROMA_STRIDE);
      }
      u32 sse = sse_8x8_chroma(ubuf.data(), chroma_width, mbx, mby, rec_u) + sse_8x8_chroma(vbuf.data(), chroma_width, mbx, mby, rec_v);
      u32 rate_penalty = 0;
      if (mode_idx > 0) {
        i16 all_levels[128];
        for (size_t b = 0; b < 8; b++) for (size_t i = 0; i < 16; i++) all_levels[b * 16 + i] = (i16)uv_quant[b][i];
        if (is_flat_coeffs(all_levels, 8, FLATNESS_LIMIT_UV)) rate_penalty = FLATNESS_PENALTY * 8;
      }
      u16 mode_cost = kFixedCostsUV[mode_idx];
      i64 rate = ((i64)mode_cost + (i64)coeff_cost + (i64)rate_penalty) * (i64)lambda;
      i64 distortion = (i64)RD_DISTO_MULT * (i64)sse;
      i64 rds = rate + distortion;
      if (rds < best_rd_score) { best_rd_score = rds; best_mode = mode; }
    }
    return best_mode;
  }

  MacroblockInfo choose_macroblock_info(size_t mbx, size_t mby) const {  // vp8.rs:2202-2246
    OpsGate ops_gate_;
    u64 i16_score;
    int luma_mode = pick_best_intra16(mbx, mby, i16_score);
    MacroblockInfo info;
    info.luma_mode = luma_mode;
    if (method > 1) {
      const Segment& segment = get_segment_for_mb(mbx, mby);
      u64 skip_i4_threshold = 211ull * (u64)segment.lambda_mode;
      bool should_try_i4 = method >= 5 || i16_score > skip_i4_threshold || luma_mode != LM_DC;
      if (should_try_i4) {
        int modes[16];
        if (pick_best_intra4(mbx, mby, i16_score, modes)) {
          info.luma_mode = LM_B;
          info.has_bpred = true;
          memcpy(info.luma_bpred, modes, sizeof(modes));
        }
      }
    }
    info.chroma_mode = pick_best_uv(mbx, mby);
    info.segment_id = get_segment_id_for_mb(mbx, mby);
    info.coeffs_skipped = false;
    return info;
  }

  // ---- final transforms -------------------------------------------------------------------
  void transform_luma_blocks_4x4(const int* bpred_modes, size_t mbx, size_t mby, i32* luma_blocks) {  // vp8.rs:2785-2916
    OpsGate ops_gate_;
    memset(luma_blocks, 0, 256 * sizeof(i32));
    size_t stride = LUMA_STRIDE, mbw = macroblock_width, w = mbw * 16;
    LumaBuf y_with_border = create_border_luma(mbx, mby, mbw, top_border_y.data(), left_border_y);
    const Segment& segment = get_segment_for_mb(mbx, mby);
    bool trellis = do_trellis;
    u32 lambda = segment.lambda_trellis_i4;
    bool top_nz[4], left_nz[4];
    for (int i = 0; i < 4; i++) { top_nz[i] = top_complexity[mbx].y[i] != 0; left_nz[i] = left_complexity.y[i] != 0; }
    for (size_t sby = 0; sby < 4; sby++)
      for (size_t sbx = 0; sbx < 4; sbx++) {
        size_t i = sby * 4 + sbx, y0 = sby * 4 + 1, x0 = sbx * 4 + 1;
        apply_intra4_prediction(y_with_border.data(), bpred_modes[i], x0, y0, stride);
        size_t block_index = sby * 16 * 4 + sbx * 16;
        i32 cur[16];
        size_t border_subblock_index = y0 * stride + x0;
        size_t y_data_block_index = (mby * 16 + sby * 4) * w + mbx * 16 + sbx * 4;
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++)
            cur[y * 4 + x] = (i32)ybuf[y_data_block_index + y * w + x] - (i32)y_with_border[border_subblock_index + y * stride + x];
        dct4x4(cur);
        memcpy(luma_blocks + block_index, cur, sizeof(cur));
        const VP8Matrix& y1_matrix = segment.y1_matrix;
        bool has_nz;
        if (trellis) {
          size_t ctx0 = std::min<u8>((u8)left_nz[sby] + (u8)top_nz[sbx], 2);
          i32 zigzag_out[16] = {0};
          has_nz = trellis_quantize_block(cur, zigzag_out, y1_matrix, lambda, 0, level_costs, 3, ctx0);
        } else {
          has_nz = false;
          for (size_t idx = 0; idx < 16; idx++) {
            i32 level = y1_matrix.quantize_coeff(cur[idx], idx);
            has_nz |= level != 0;
            cur[idx] = y1_matrix.dequantize(level, idx);
          }
        }
        top_nz[sbx] = has_nz;
        left_nz[sby] = has_nz;
        idct4x4(cur);
        add_residue(y_with_border.data(), cur, y0, x0, stride);
      }
    for (size_t y = 0; y < 17; y++) left_border_y[y] = y_with_border[y * stride + 16];
    for (size_t x = 0; x < 16; x++) top_border_y[mbx * 16 + x] = y_with_border[16 * stride + x + 1];
  }

  void transform_luma_block(size_t mbx, size_t mby, const MacroblockInfo& info, i32* luma_blocks) {  // vp8.rs:2647-2780
    OpsGate ops_gate_;
    if (info.luma_mode == LM_B) {
      assert(info.has_bpred);
      transform_luma_blocks_4x4(info.luma_bpred, mbx, mby, luma_blocks);
      return;
    }
    LumaBuf y_with_border = get_predicted_luma_block_16x16(info.luma_mode, mbx, mby);
    get_luma_blocks_from_predicted_16x16(y_with_border, mbx, mby, luma_blocks);
    const Segment& segment = segments[info.segment_id < 0 ? 0 : info.segment_id];
    const VP8Matrix &y1_matrix = segment.y1_matrix, &y2_matrix = segment.y2_matrix;
    i32 coeffs0[16];
    get_coeffs0_from_block(luma_blocks, coeffs0);
    wht4x4(coeffs0);
    for (size_t k = 0; k < 16; k++) coeffs0[k] = y2_matrix.quantize_coeff(coeffs0[k], k);
    bool trellis = do_trellis;
    u32 lambda = segment.lambda_trellis_i16;
    bool top_nz[4], left_nz[4];
    for (int i = 0; i < 4; i++) { top_nz[i] = top_complexity[mbx].y[i] != 0; left_nz[i] = left_complexity.y[i] != 0; }
    i32 y2_dequant[16];
    for (size_t k = 0; k < 16; k++) y2_dequant[k] = y2_matrix.dequantize(coeffs0[k], k);
    iwht4x4(y2_dequant);
    i32 dequantized_blocks[256];
    for (size_t y = 0; y < 4; y++)
      for (size_t x = 0; x < 4; x++) {
        size_t i = y * 4 + x;
        i32 block[16];
        memcpy(block, luma_blocks + i * 16, sizeof(block));
        if (trellis) {
          size_t ctx0 = std::min<u8>((u8)left_nz[y] + (u8)top_nz[x], 2);
          i32 zigzag_out[16] = {0};
          bool has_nz = trellis_quantize_block(block, zigzag_out, y1_matrix, lambda, 1, level_costs, 0, ctx0);
          top_nz[x] = has_nz;
          left_nz[y] = has_nz;
        } else {
          bool has_nz = false;
          for (size_t idx = 0; idx < 16; idx++) {
            if (idx == 0) { block[idx] = 0; }
            else {
              i32 level = y1_matrix.quantize_coeff(block[idx], idx);
              has_nz |= level != 0;
              block[idx] = y1_matrix.dequantize(level, idx);
            }
          }
          top_nz[x] = has_nz;
          left_nz[y] = has_nz;
        }
        block[0] = y2_dequant[i];
        idct4x4(block);
        memcpy(dequantized_blocks + i * 16, block, sizeof(block));
      }
    for (size_t y = 0; y < 4; y++)
      for (size_t x = 0; x < 4; x++) {
        size_t i = x + y * 4;
        add_residue(y_with_border.data(), dequantized_blocks + i * 16, 1 + y * 4, 1 + x * 4, LUMA_STRIDE);
      }
    for (size_t y = 0; y < 17; y++) left_border_y[y] = y_with_border[y * LUMA_STRIDE + 16];
    for (size_t x = 0; x < 16; x++) top_border_y[mbx * 16 + x] = y_with_border[16 * LUMA_STRIDE + x + 1];
  }

  void transform_chroma_blocks(size_t mbx, size_t mby, int chroma_mode, i32* u_blocks, i32* v_blocks) {  // vp8.rs:3039-3121
    OpsGate ops_gate_;
    OpsChroma ops_chroma_;
    size_t stride = CHROMA_STRIDE;
    ChromaBuf predicted_u = get_predicted_chroma_block(chroma_mode, mbx, mby, top_border_u, left_border_u);
    ChromaBuf predicted_v = get_predicted_chroma_block(chroma_mode, mbx, mby, top_border_v, left_border_v);
    get_chroma_blocks_from_predicted(predicted_u, ubuf, mbx, mby, u_blocks);
    get_chroma_blocks_from_predicted(predicted_v, vbuf, mbx, mby, v_blocks);
    const Segment& segment = get_segment_for_mb(mbx, mby);
    VP8Matrix uv_matrix = segment.uv_matrix;
    if (do_error_diffusion) apply_chroma_error_diffusion(u_blocks, v_blocks, mbx, uv_matrix);
    i32 u_res[64], v_res[64];
    for (size_t b = 0; b < 4; b++) {
      for (size_t i = 0; i < 16; i++) {
        u_res[b * 16 + i] = uv_matrix.dequantize(uv_matrix.quantize_coeff(u_blocks[b * 16 + i], i), i);
        v_res[b * 16 + i] = uv_matrix.dequantize(uv_matrix.quantize_coeff(v_blocks[b * 16 + i], i), i);
      }
      idct4x4(u_res + b * 16);
      idct4x4(v_res + b * 16);
    }
    for (size_t y = 0; y < 2; y++)
      for (size_t x = 0; x < 2; x++) {
        size_t i = x + y * 2;
        add_residue(predicted_u.data(), u_res + i * 16, 1 + y * 4, 1 + x * 4, stride);
        add_residue(predicted_v.data(), v_res + i * 16, 1 + y * 4, 1 + x * 4, stride);
      }
    for (size_t y = 0; y < 9; y++) { left_border_u[y] = predicted_u[y * stride + 8]; left_border_v[y] = predicted_v[y * stride + 8]; }
    for (size_t x = 0; x < 8; x++) { top_border_u[mbx * 8 + x] = predicted_u[8 * stride + x + 1]; top_border_v[mbx * 8 + x] = predicted_v[8 * stride + x + 1]; }
  }

  // ---- setup ------------------------------------------------------------------------------
  void analyze_and_assign_segments(u8 base_quant_index) {  // vp8.rs:2278-2388
    size_t y_stride = (size_t)macroblock_width * 16, uv_stride = (size_t)macroblock_width * 8;
    std::vector<u8> mb_alphas;
    u32 alpha_histogram[256];
    analyze_image(ybuf.data(), ubuf.data(), vbuf.data(), width, height, y_stride, uv_stride, mb_alphas, alpha_histogram);
    u8 centers[4], alpha_to_segment[256];
    i32 mid_alpha;
    assign_segments_kmeans(alpha_histogram, 4, centers, alpha_to_segment, mid_alpha);
    i32 min_center = *std::min_element(centers, centers + 4), max_center = *std::max_element(centers, centers + 4);
    i32 range = max_center == min_center ? 1 : max_center - min_center;
    segment_map.resize(mb_alphas.size());
    for (size_t i = 0; i < mb_alphas.size(); i++) segment_map[i] = alpha_to_segment[mb_alphas[i]];
    const u8 sns_strength = 50;
    u8 qidx[4];
    for (size_t s = 0; s < 4; s++) {
      i32 center = centers[s];
      i32 transformed_alpha = std::min(std::max(255 * (center - mid_alpha) / range, -127), 127);
      u8 seg_quant_index = compute_segment_quant(base_quant_index, transformed_alpha, sns_strength);
      i8 delta = (i8)((i8)seg_quant_index - (i8)base_quant_index);
      segments[s] = make_segment(seg_quant_index, delta);
      qidx[s] = seg_quant_index;
    }
    u32 seg_counts[4] = {0, 0, 0, 0};
    for (u8 id : segment_map) seg_counts[id] += 1;
    auto get_proba = [](u32 a, u32 b) -> u8 {
      u32 total = a + b;
      if (total == 0) return 255;
      return (u8)((255 * a + total / 2) / total);
    };
    segment_tree_probs[0] = get_proba(seg_counts[0] + seg_counts[1], seg_counts[2] + seg_counts[3]);
    segment_tree_probs[1] = get_proba(seg_counts[0], seg_counts[1]);
    segment_tree_probs[2] = get_proba(seg_counts[2], seg_counts[3]);
    bool should_update_map = segment_tree_probs[0] != 255 || segment_tree_probs[1] != 255 || segment_tree_probs[2] != 255;
    segments_enabled = true;
    segments_update_map = should_update_map;
    memset(left_border_y, 129, 17); memset(left_border_u, 129, 9); memset(left_border_v, 129, 9);
    if (dump) {
      dump->put("ALPHA", mb_alphas.data(), mb_alphas.size());
      dump->put("ALPHA_HIST", alpha_histogram, 256);
      dump->put("SEG_CENTERS", centers, 4);
      dump->put("SEG_MAP256", alpha_to_segment, 256);
      dump->put("SEG_MID", &mid_alpha, 1);
      dump->put("SEG_MAP", segment_map.data(), segment_map.size());
      dump->put("SEG_QIDX", qidx, 4);
      dump->put("SEG_TREE_PROBS", segment_tree_probs, 3);
      u8 um = segments_update_map;
      dump->put("SEG_UPDATE_MAP", &um, 1);
    }
  }

  void setup_encoding(u8 lossy_quality, u16 w, u16 h) {  // vp8.rs:2391-2506 (planes already stored)
    u8 quant_index = quality_to_quant_index(lossy_quality);
    macroblock_width = (u16)((w + 15) / 16);
    macroblock_height = (u16)((h + 15) / 16);
    filter_level = compute_filter_level(quant_index, 0, 50);
    width = w; height = h;
    filter_type = false; sharpness_level = 0;
    top_complexity.assign(macroblock_width, Complexity());
    top_b_pred.assign(4 * (size_t)macroblock_width, IM_DC);
    for (int i = 0; i < 4; i++) left_b_pred[i] = IM_DC;
    memcpy(token_probs, kCoeffProbs, sizeof(token_probs));
    has_no_skip = true;
    macroblock_no_skip_coeff = 200;
    yac_abs = quant_index;
    for (int s = 0; s < 4; s++) segments[s] = make_segment(quant_index, 0);
    size_t total_mbs = (size_t)macroblock_width * (size_t)macroblock_height;
    bool use_segments = total_mbs >= 256;
    if (use_segments) {
      analyze_and_assign_segments(quant_index);
    } else {
      segments_enabled = false;
      segments_update_map = false;
      segment_map.clear();
    }
    memset(left_border_y, 129, 17); memset(left_border_u, 129, 9); memset(left_border_v, 129, 9);
    top_border_y.assign((size_t)macroblock_width * 16 + 4, 127);
    top_border_u.assign((size_t)macroblock_width * 8, 127);
    top_border_v.assign((size_t)macroblock_width * 8, 127);
    top_derr.assign(macroblock_width, {{{0, 0}, {0, 0}}});
    memset(left_derr, 0, sizeof(left_derr));
    if (dump) {
      u8 qi = quant_index;
      dump->put("BASE_QIDX", &qi, 1);
      u8 se = segments_enabled;
      dump->put("SEG_ENABLED", &se, 1);
    }
  }

  void fill_record(zwo_mb_record& r, const MacroblockInfo& info, size_t mbx, const Complexity& top_in, const Complexity& left_in) {
    r.ymode = (u8)info.luma_mode;
    r.uvmode = (u8)info.chroma_mode;
    r.segment = (u8)(info.segment_id < 0 ? 0 : info.segment_id);
    for (int i = 0; i < 16; i++) r.bmodes[i] = info.luma_mode == LM_B ? (u8)info.luma_bpred[i] : 0;
    r.top_nz = top_in.pack();
    r.left_nz = left_in.pack();
    r.derr_left[0] = left_derr[0][0]; r.derr_left[1] = left_derr[0][1]; r.derr_left[2] = left_derr[1][0]; r.derr_left[3] = left_derr[1][1];
    r.derr_top[0] = top_derr[mbx][0][0]; r.derr_top[1] = top_derr[mbx][0][1]; r.derr_top[2] = top_derr[mbx][1][0]; r.derr_top[3] = top_derr[mbx][1][1];
  }

  int encode_image(const u8* data, size_t data_len, int color, u16 w, u16 h, u8 lossy_quality, u8 meth) {  // vp8.rs:1281-1488
    method = std::min<u8>(meth, 6);
    do_trellis = method >= 4;
    size_t bpp = color == 0 ? 1 : color == 1 ? 2 : color == 2 ? 3 : 4;
    if ((u64)w * (u64)h * bpp != (u64)data_len) return 2;  // assert_eq! panic in the reference (:1307)
    if (lossy_quality > 100) return 3;                     // panic in the reference (:2401)
    if (w == 0 || h == 0) return 1;                        // reference would index out of bounds
    if (color == 2 || color == 3) convert_image_yuv(data, bpp, w, h, ybuf, ubuf, vbuf);
    else convert_image_y(data, bpp, w, h, ybuf, ubuf, vbuf);
    if (dump) { dump->put("YUV_Y", ybuf.data(), ybuf.size()); dump->put("YUV_U", ubuf.data(), ubuf.size()); dump->put("YUV_V", vbuf.data(), vbuf.size()); }
    setup_encoding(lossy_quality, w, h);
    size_t nmb = (size_t)macroblock_width * macroblock_height;
    if (dump) { rec_p1.assign(nmb, zwo_mb_record()); rec_p2.assign(nmb, zwo_mb_record()); }

    // ===== PASS 1 =====
    g_ops_pass = 0;
    {
      bool trellis_during_pass2 = do_trellis;
      do_trellis = false;
      proba_stats.reset();
      u32 total_mb = 0, skip_mb = 0;
      for (size_t mby = 0; mby < macroblock_height; mby++) {
        left_complexity = Complexity();
        for (int i = 0; i < 4; i++) left_b_pred[i] = IM_DC;
        memset(left_border_y, 129, 17); memset(left_border_u, 129, 9); memset(left_border_v, 129, 9);
        for (size_t mbx = 0; mbx < macroblock_width; mbx++) {
          MacroblockInfo info = choose_macroblock_info(mbx, mby);
          i32 y_block_data[256], u_block_data[64], v_block_data[64];
          transform_luma_block(mbx, mby, info, y_block_data);
          transform_chroma_blocks(mbx, mby, info.chroma_mode, u_block_data, v_block_data);
          zwo_mb_record* rec = dump ? &rec_p1[mby * macroblock_width + mbx] : nullptr;
          if (rec) { memset(rec, 0, sizeof(*rec)); fill_record(*rec, info, mbx, top_complexity[mbx], left_complexity); }
          total_mb += 1;
          bool all_zero = check_all_coeffs_zero(info, y_block_data, u_block_data, v_block_data);
          if (all_zero) {
            skip_mb += 1;
            left_complexity.clear(info.luma_mode != LM_B);
            top_complexity[mbx].clear(info.luma_mode != LM_B);
          } else {
            record_residual_stats(info, mbx, y_block_data, u_block_data, v_block_data, rec);
          }
          if (rec) rec->skip = all_zero;
        }
      }
      if (total_mb > 0) {
        u32 non_skip_mb = total_mb - skip_mb;
        u8 prob = (u8)std::min<u32>((255 * non_skip_mb + total_mb / 2) / total_mb, 255);
        has_no_skip = true;
        macroblock_no_skip_coeff = std::min<u8>(std::max<u8>(prob, 1), 254);
      }
      compute_updated_probabilities();
      level_costs.calculate(cur_probs());
      do_trellis = trellis_during_pass2;
      reset_for_second_pass();
      if (dump) {
        dump->put("P1MB", rec_p1.data(), rec_p1.size());
        dump->put("STATS", &proba_stats.stats[0][0][0][0], 1056);
        dump->put("PROBS", &cur_probs()[0][0][0][0], 1056);
        u8 hu = has_updated_probs;
        dump->put("PROBS_UPDATED", &hu, 1);
        dump->put("SKIP_PROB", &macroblock_no_skip_coeff, 1);
        dump->put("LCOST", &level_costs.level_cost[0][0][0][0], 4 * 8 * 3 * 68);
      }
    }
    if (level_costs.dirty) level_costs.calculate(token_probs);

    std::vector<u16> sym_hdr, sym_tok;  // dump only: the symbol streams of the two boolean coders
    if (dump) { encoder.log = &sym_hdr; partitions[0].log = &sym_tok; }
    encode_compressed_frame_header();

    // ===== PASS 2 =====
    g_ops_pass = 1;
    for (size_t mby = 0; mby < macroblock_height; mby++) {
      size_t partition_index = mby % partitions.size();
      left_complexity = Complexity();
      for (int i = 0; i < 4; i++) left_b_pred[i] = IM_DC;
      memset(left_derr, 0, sizeof(left_derr));
      memset(left_border_y, 129, 17); memset(left_border_u, 129, 9); memset(left_border_v, 129, 9);
      for (size_t mbx = 0; mbx < macroblock_width; mbx++) {
        MacroblockInfo info = choose_macroblock_info(mbx, mby);
        i32 y_block_data[256], u_block_data[64], v_block_data[64];
        transform_luma_block(mbx, mby, info, y_block_data);
        transform_chroma_blocks(mbx, mby, info.chroma_mode, u_block_data, v_block_data);
        if (has_no_skip) info.coeffs_skipped = check_all_coeffs_zero(info, y_block_data, u_block_data, v_block_data);
        zwo_mb_record* rec = dump ? &rec_p2[mby * macroblock_width + mbx] : nullptr;
        if (rec) { memset(rec, 0, sizeof(*rec)); fill_record(*rec, info, mbx, top_complexity[mbx], left_complexity); rec->skip = info.coeffs_skipped; }
        write_macroblock_header(info, mbx);
        if (!info.coeffs_skipped) {
          encode_residual_data(info, partition_index, mbx, y_block_data, u_block_data, v_block_data, rec);
        } else {
          left_complexity.clear(info.luma_mode != LM_B);
          top_complexity[mbx].clear(info.luma_mode != LM_B);
        }
      }
    }
    std::vector<u8> compressed_header_bytes = encoder.flush_and_get_buffer();
    encoder = ArithmeticEncoder();
    if (dump) {
      dump->put("HDR_TOKENS", sym_hdr.data(), sym_hdr.size());
      dump->put("TOK_TOKENS", sym_tok.data(), sym_tok.size());
      dump->put("P2MB", rec_p2.data(), rec_p2.size());
      dump->put("PART0", compressed_header_bytes.data(), compressed_header_bytes.size());
      dump->put("TOKEN_PROBS_FINAL", &token_probs[0][0][0][0], 1056);
    }
    write_uncompressed_frame_header((u32)compressed_header_bytes.size());
    writer->insert(writer->end(), compressed_header_bytes.begin(), compressed_header_bytes.end());
    write_partitions();
    return 0;
  }
};

static int encode_frame_lossy(std::vector<u8>& writer, const u8* data, size_t data_len, u32 width, u32 height, int color,
                              int quality, int method, zwo_dump* dump) {  // vp8.rs:3132-3153
  if (width > 65535 || height > 65535) return 1;  // InvalidDimensions
  if (quality < 0 || quality > 255 || method < 0 || method > 255) return 3;
  Vp8Encoder enc(&writer);
  enc.dump = dump;
  return enc.encode_image(data, data_len, color, (u16)width, (u16)height, (u8)quality, (u8)method);
}

// src/encoder/api.rs:1224-1241, 1320-1329
static u32 chunk_size(size_t inner) { return (inner % 2 == 1) ? (u32)(inner + 1) + 8 : (u32)inner + 8; }
static void write_u32_le(std::vector<u8>& w, u32 v) { for (int i = 0; i < 4; i++) w.push_back((u8)(v >> (8 * i))); }
static void write_chunk(std::vector<u8>& w, const char* name, const std::vector<u8>& data) {
  w.insert(w.end(), name, name + 4);
  write_u32_le(w, (u32)data.size());
  w.insert(w.end(), data.begin(), data.end());
  if (data.size() % 2 == 1) w.push_back(0);
}

#include "zw_dec_oracle.inc"
#include "zw_lossless_oracle.inc"

}  // namespace

// =============================================================================================
// C interface
// =============================================================================================
extern "C" {

int zwo_encode_vp8(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height, int color, int quality,
                   int method, uint8_t** out, size_t* out_len, zwo_dump* dump) {
  std::vector<u8> w;
  int rc = encode_frame_lossy(w, data, data_len, width, height, color, quality, method, dump);
  if (rc != 0) { *out = nullptr; *out_len = 0; return rc; }
  if (dump) dump->put("VP8", w.data(), w.size());
  *out = (uint8_t*)malloc(w.size() ? w.size() : 1);
  memcpy(*out, w.data(), w.size());
  *out_len = w.size();
  return 0;
}

int zwo_encode_webp(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height, int color, int quality,
                    int method, uint8_t** out, size_t* out_len, zwo_dump* dump) {
  std::vector<u8> frame;
  int rc = encode_frame_lossy(frame, data, data_len, width, height, color, quality, method, dump);
  if (rc != 0) { *out = nullptr; *out_len = 0; return rc; }
  if (dump) dump->put("VP8", frame.data(), frame.size());
  std::vector<u8> w;
  const char* riff = "RIFF";
  w.insert(w.end(), riff, riff + 4);
  write_u32_le(w, chunk_size(frame.size()) + 4);
  const char* webp = "WEBP";
  w.insert(w.end(), webp, webp + 4);
  write_chunk(w, "VP8 ", frame);
  *out = (uint8_t*)malloc(w.size());
  memcpy(*out, w.data(), w.size());
  *out_len = w.size();
  return 0;
}


// ---- decoder (zw_dec_oracle.inc): test infrastructure for the on-device decoder / verifier ----
// Decodes a VP8 key frame (bare, or inside RIFF/WEBP).  Any output pointer may be NULL.  *rgb: width*height*3
// (fancy != 0: bilinear chroma upsampling like WebPDecoder::read_image's default, else nearest); *planes: Y | U | V of
// the padded frame (16*mbw x 16*mbh, 8*mbw x 8*mbh twice) after the loop filter, *planes_unfiltered the same before it;
// *mbinfo: 24 bytes per macroblock {luma_mode, chroma_mode, segment, coeffs_skipped, non_zero_dct, 0,0,0, bpred[16]};
// hdr[16]: width, height, mbw, mbh, filter_type, filter_level, sharpness, num_partitions, segments_enabled, update_map,
// lf_adj, has_skip_prob, prob_skip_false, version, pixel_type, first-partition bytes consumed (approx.)
int zwo_decode(const uint8_t* data, size_t len, int fancy, uint8_t** rgb, uint8_t** planes, uint8_t** planes_unfiltered,
               uint8_t** mbinfo, uint32_t hdr[16]) {
  if (rgb) *rgb = nullptr;
  if (planes) *planes = nullptr;
  if (planes_unfiltered) *planes_unfiltered = nullptr;
  if (mbinfo) *mbinfo = nullptr;
  const u8* vp8 = nullptr;
  size_t vlen = 0;
  int rc = dy_find_vp8(data, len, &vp8, &vlen);
  if (rc != ZWD_OK) return rc;
  Vp8DecoderO d;
  d.data = vp8; d.len = vlen;
  d.keep_unfiltered = planes_unfiltered != nullptr;
  rc = d.decode_frame();
  if (hdr) {
    const uint32_t h[16] = {d.width, d.height, d.mbwidth, d.mbheight, d.filter_type, d.filter_level, d.sharpness_level, (uint32_t)d.num_partitions,
                            d.segments_enabled, d.segments_update_map, d.lf_adj, d.has_skip_prob, d.prob_skip_false, d.version, d.pixel_type, 0};
    memcpy(hdr, h, sizeof(h));
  }
  if (rc != ZWD_OK) return rc;
  const size_t ysz = d.ybuf.size(), csz = d.ubuf.size();
  if (rgb) {
    *rgb = (uint8_t*)malloc((size_t)d.width * d.height * 3 + 1);
    dy_fill_rgb(*rgb, d, fancy != 0);
  }
  if (planes) {
    *planes = (uint8_t*)malloc(ysz + 2 * csz);
    memcpy(*planes, d.ybuf.data(), ysz); memcpy(*planes + ysz, d.ubuf.data(), csz); memcpy(*planes + ysz + csz, d.vbuf.data(), csz);
  }
  if (planes_unfiltered) {
    *planes_unfiltered = (uint8_t*)malloc(ysz + 2 * csz);
    memcpy(*planes_unfiltered, d.ybuf0.data(), ysz); memcpy(*planes_unfiltered + ysz, d.ubuf0.data(), csz); memcpy(*planes_unfiltered + ysz + csz, d.vbuf0.data(), csz);
  }
  if (mbinfo) {
    const size_t nmb = d.macroblocks.size();
    *mbinfo = (uint8_t*)calloc(nmb ? nmb : 1, 24);
    for (size_t i = 0; i < nmb; i++) {
      const DecMacroBlock& m = d.macroblocks[i];
      uint8_t* o = *mbinfo + i * 24;
      o[0] = m.luma_mode; o[1] = m.chroma_mode; o[2] = m.segmentid; o[3] = m.coeffs_skipped; o[4] = m.non_zero_dct;
      memcpy(o + 8, m.bpred, 16);
    }
  }
  return ZWD_OK;
}

// ---- lossless VP8L encoder + full container (zw_lossless_oracle.inc) ----
static int ll_finish(int rc, std::vector<u8>& w, uint8_t** out, size_t* out_len) {
  if (rc != 0) { *out = nullptr; *out_len = 0; return rc; }
  *out = (uint8_t*)malloc(w.size() ? w.size() : 1);
  memcpy(*out, w.data(), w.size());
  *out_len = w.size();
  return 0;
}
int zwo_encode_lossless(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height, int color, int use_predictor,
                        int implicit_dimensions, uint8_t** out, size_t* out_len) {
  std::vector<u8> w;
  return ll_finish(ll::encode_frame_lossless(w, data, data_len, width, height, color, use_predictor != 0, implicit_dimensions != 0), w, out, out_len);
}
int zwo_encode_alpha_lossless(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height, int color, uint8_t** out,
                              size_t* out_len) {
  std::vector<u8> w;
  return ll_finish(ll::encode_alpha_lossless(w, data, data_len, width, height, color), w, out, out_len);
}
int zwo_webp_encode(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height, int color, int use_predictor,
                    int use_lossy, int quality, int method, const uint8_t* icc, size_t icc_len, const uint8_t* exif,
                    size_t exif_len, const uint8_t* xmp, size_t xmp_len, uint8_t** out, size_t* out_len) {
  std::vector<u8> w;
  const ll::Meta m = {icc, icc_len, exif, exif_len, xmp, xmp_len};
  return ll_finish(ll::webp_encode(w, data, data_len, width, height, color, use_predictor != 0, use_lossy != 0, quality, method, m), w, out, out_len);
}
int zwo_build_huffman(const uint32_t* frequencies, size_t n, int length_limit, uint8_t* lengths, uint16_t* codes) {
  return ll::build_huffman_tree(frequencies, n, lengths, codes, (u8)length_limit) ? 1 : 0;
}

// WebPEncoder::encode (lossy or lossless, no metadata) for n same-sized images on `threads` host threads; outputs as in
// zwo_encode_batch_mt_out.  Used by bench.py's lossless CPU baseline and parity check.
size_t zwo_webp_encode_batch_mt(const uint8_t* data, size_t n, uint32_t width, uint32_t height, int color, int use_predictor,
                                int use_lossy, int quality, int method, int threads, uint8_t* out, size_t out_stride,
                                uint32_t* out_lens) {
  std::atomic<size_t> next(0), total(0);
  const size_t per = (size_t)width * height * (size_t)(color + 1);
  auto work = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) break;
      uint8_t* o = nullptr;
      size_t len = 0;
      const int rc = zwo_webp_encode(data + i * per, per, width, height, color, use_predictor, use_lossy, quality, method, nullptr, 0,
                                     nullptr, 0, nullptr, 0, &o, &len);
      if (rc != 0) { if (out_lens) out_lens[i] = 0; continue; }
      if (out) memcpy(out + i * out_stride, o, len < out_stride ? len : out_stride);
      if (out_lens) out_lens[i] = (uint32_t)len;
      total.fetch_add(len);
      free(o);
    }
  };
  if (threads <= 1) { work(); return total.load(); }
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; t++) ts.emplace_back(work);
  for (auto& t : ts) t.join();
  return total.load();
}

void zwo_free(void* p) { free(p); }
// Counters of the calling thread: reset, then run zwo_encode_* on this thread, then read.
void zwo_opcounts_reset(void) { memset(g_ops, 0, sizeof(g_ops)); }
size_t zwo_opcounts_get(uint64_t* out, size_t cap) {
  for (size_t i = 0; i < (size_t)(4 * OPC_N) && i < cap; i++) out[i] = g_ops[i / OPC_N][i % OPC_N];
  return (size_t)OPC_N;
}

size_t zwo_encode_batch_mt(const uint8_t* data, size_t n, uint32_t width, uint32_t height, int quality, int method,
                           int threads) {
  std::atomic<size_t> next(0), total(0);
  size_t per = (size_t)width * height * 3;
  auto work = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) break;
      std::vector<u8> frame;
      encode_frame_lossy(frame, data + i * per, per, width, height, 2, quality, method, nullptr);
      total.fetch_add(frame.size());
    }
  };
  if (threads <= 1) { work(); return total.load(); }
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; t++) ts.emplace_back(work);
  for (auto& t : ts) t.join();
  return total.load();
}

// Same, keeping the outputs: image i's file (container != 0: RIFF-wrapped) is copied to out + i * out_stride
// (truncated there if longer; out_lens[i] always holds the full length, 0 on failure).
size_t zwo_encode_batch_mt_out(const uint8_t* data, size_t n, uint32_t width, uint32_t height, int color, int quality, int method,
                               int threads, int container, uint8_t* out, size_t out_stride, uint32_t* out_lens) {
  std::atomic<size_t> next(0), total(0);
  const size_t per = (size_t)width * height * (size_t)(color + 1);
  auto work = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) break;
      uint8_t* o = nullptr;
      size_t len = 0;
      const int rc = container ? zwo_encode_webp(data + i * per, per, width, height, color, quality, method, &o, &len, nullptr)
                               : zwo_encode_vp8(data + i * per, per, width, height, color, quality, method, &o, &len, nullptr);
      if (rc != 0) { if (out_lens) out_lens[i] = 0; continue; }
      if (out) memcpy(out + i * out_stride, o, len < out_stride ? len : out_stride);
      if (out_lens) out_lens[i] = (uint32_t)len;
      total.fetch_add(len);
      free(o);
    }
  };
  if (threads <= 1) { work(); return total.load(); }
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; t++) ts.emplace_back(work);
  for (auto& t : ts) t.join();
  return total.load();
}

zwo_dump* zwo_dump_new(void) { return new zwo_dump(); }
void zwo_dump_free(zwo_dump* d) { delete d; }
int zwo_dump_get(const zwo_dump* d, const char* name, const uint8_t** ptr, size_t* len) {
  auto it = d->stages.find(name);
  if (it == d->stages.end()) return 0;
  *ptr = it->second.data();
  *len = it->second.size();
  return 1;
}

void zwo_dct4x4(int32_t b[16]) { dct4x4(b); }
void zwo_idct4x4(int32_t b[16]) { idct4x4(b); }
void zwo_wht4x4(int32_t b[16]) { wht4x4(b); }
void zwo_iwht4x4(int32_t b[16]) { iwht4x4(b); }

size_t zwo_bool_encode(const uint8_t* bits, const uint8_t* probs, size_t n, uint8_t* out) {
  ArithmeticEncoder e;
  for (size_t i = 0; i < n; i++) e.write_bool(bits[i] != 0, probs[i]);
  std::vector<u8> b = e.flush_and_get_buffer();
  memcpy(out, b.data(), b.size());
  return b.size();
}

size_t zwo_bool_encode_tree(int tree_id, const uint8_t* probs, const int8_t* values, size_t n, int start_index,
                            uint8_t* out) {
  const i8* tree; size_t len;
  switch (tree_id) {
    case 0: tree = DCT_TOKEN_TREE; len = 22; break;
    case 1: tree = KEYFRAME_YMODE_TREE; len = 8; break;
    case 2: tree = KEYFRAME_BPRED_MODE_TREE; len = 18; break;
    case 3: tree = KEYFRAME_UV_MODE_TREE; len = 6; break;
    default: tree = SEGMENT_ID_TREE; len = 6; break;
  }
  ArithmeticEncoder e;
  for (size_t i = 0; i < n; i++) e.write_with_tree_start_index(tree, len, probs, values[i], (size_t)start_index);
  std::vector<u8> b = e.flush_and_get_buffer();
  memcpy(out, b.data(), b.size());
  return b.size();
}

int zwo_trellis(int32_t coeffs[16], int32_t out[16], const uint16_t q[16], const uint32_t iq[16],
                const uint32_t bias[16], const uint16_t sharpen[16], uint32_t lambda, int first,
                const uint8_t* probs1056, int ctype, int ctx0) {
  VP8Matrix m;
  memset(&m, 0, sizeof(m));
  for (int i = 0; i < 16; i++) { m.q[i] = q[i]; m.iq[i] = iq[i]; m.bias[i] = bias[i]; m.sharpen[i] = sharpen[i]; }
  LevelCosts lc;
  TokenProbTables p;
  memcpy(p, probs1056, sizeof(p));
  lc.calculate(p);
  return trellis_quantize_block(coeffs, out, m, lambda, (size_t)first, lc, (size_t)ctype, (size_t)ctx0) ? 1 : 0;
}

void zwo_matrix_new(int q_dc, int q_ac, int type, uint16_t q[16], uint32_t iq[16], uint32_t bias[16],
                    uint32_t zthresh[16], uint16_t sharpen[16]) {
  VP8Matrix m = VP8Matrix::make((u16)q_dc, (u16)q_ac, (MatrixType)type);
  memcpy(q, m.q, sizeof(m.q)); memcpy(iq, m.iq, sizeof(m.iq)); memcpy(bias, m.bias, sizeof(m.bias));
  memcpy(zthresh, m.zthresh, sizeof(m.zthresh)); memcpy(sharpen, m.sharpen, sizeof(m.sharpen));
}

int zwo_quality_to_quant_index(int quality) { return quality_to_quant_index((u8)quality); }
void zwo_segment_lambdas(int quant_index, uint32_t l[8], int16_t quants[6]) {
  Segment s = make_segment((u8)quant_index, 0);
  l[0] = s.lambda_i4; l[1] = s.lambda_i16; l[2] = s.lambda_uv; l[3] = s.lambda_mode;
  l[4] = s.lambda_trellis_i4; l[5] = s.lambda_trellis_i16; l[6] = s.lambda_trellis_uv; l[7] = s.tlambda;
  quants[0] = s.ydc; quants[1] = s.yac; quants[2] = s.y2dc; quants[3] = s.y2ac; quants[4] = s.uvdc; quants[5] = s.uvac;
}
int zwo_compute_segment_quant(int base_quant, int segment_alpha, int sns) { return compute_segment_quant((u8)base_quant, segment_alpha, (u8)sns); }
int zwo_compute_filter_level(int qi, int sharp, int fs) { return compute_filter_level((u8)qi, (u8)sharp, (u8)fs); }
double zwo_cbrt(double x) { return fm_cbrt(x); }
double zwo_pow(double x, double n) { return fm_pow(x, n); }

void zwo_predict4x4(uint8_t* ws, int mode, int x0, int y0, int stride) { apply_intra4_prediction(ws, mode, (size_t)x0, (size_t)y0, (size_t)stride); }
void zwo_predict4x4_all(const uint8_t* ws, int x0, int y0, int stride, uint8_t out[160]) {
  I4Predictions p = i4_predictions_compute(ws, (size_t)x0, (size_t)y0, (size_t)stride);
  memcpy(out, p.data, 160);
}
void zwo_add_residue(uint8_t* pblock, const int32_t rblock[16], int y0, int x0, int stride) { add_residue(pblock, rblock, (size_t)y0, (size_t)x0, (size_t)stride); }
uint32_t zwo_rd_score(uint32_t sse, uint16_t mode_cost, uint32_t lambda, uint64_t* full) {
  u64 r = rd_score(sse, mode_cost, lambda);
  if (full) *full = r;
  return (u32)r;
}
void zwo_record_coeffs(const int32_t coeffs[16], int token_type, int first, int ctx, uint32_t stats[1056]) {
  ProbaStats s;
  memcpy(s.stats, stats, sizeof(s.stats));
  record_coeffs(coeffs, token_type, (size_t)first, (size_t)ctx, s);
  memcpy(stats, s.stats, sizeof(s.stats));
}
void zwo_level_costs(const uint8_t* probs1056, uint16_t* out) {
  LevelCosts lc;
  TokenProbTables p;
  memcpy(p, probs1056, sizeof(p));
  lc.calculate(p);
  memcpy(out, lc.level_cost, sizeof(lc.level_cost));
}
uint32_t zwo_residual_cost(const int32_t levels[16], int ctype, int first, int ctx0, const uint8_t* probs1056, int zero_tables) {
  LevelCosts lc;
  TokenProbTables p;
  memcpy(p, probs1056, sizeof(p));
  if (!zero_tables) lc.calculate(p);
  return get_residual_cost((size_t)ctx0, levels, (size_t)ctype, (size_t)first, lc, p);
}
int zwo_tdisto_16x16(const uint8_t* a, const uint8_t* b, int stride) { return tdisto_16x16(a, b, (size_t)stride, kWeightY); }
void zwo_convert_yuv(const uint8_t* rgb, uint32_t w, uint32_t h, int bpp, uint8_t* y, uint8_t* u, uint8_t* v) {
  std::vector<u8> yb, ub, vb;
  convert_image_yuv(rgb, (size_t)bpp, w, h, yb, ub, vb);
  memcpy(y, yb.data(), yb.size()); memcpy(u, ub.data(), ub.size()); memcpy(v, vb.data(), vb.size());
}
uint16_t zwo_fixed_cost_i16(int i) { return kFixedCostsI16[i]; }
uint16_t zwo_fixed_cost_uv(int i) { return kFixedCostsUV[i]; }
uint16_t zwo_fixed_cost_i4(int top, int left, int mode) { return get_i4_mode_cost((size_t)top, (size_t)left, (size_t)mode); }
uint16_t zwo_entropy_cost(int p) { return kEntropyCost[p]; }
uint16_t zwo_level_fixed_cost(int level) { return kLevelFixedCosts[level]; }

}  // extern "C"
