// zw_hostcheck.cpp -- TEST INFRASTRUCTURE: compiles the per-lane device primitives
// (image_webp_b200/csrc/zw_prims.cuh, zw_cost.cuh) with the HOST compiler so that the CPU test
// suite can compare them with the oracle on random inputs without a GPU.  Not part of the
// product; the shipped library only runs these functions inside CUDA kernels.
#include <cstring>
#include "../../image_webp_b200/csrc/zw_cost.cuh"
using namespace zw;
static const u16 kPredTab[8][16] = ZW_PRED_TABLE_INIT;
static const u16 kDTaps[32] = ZW_DTAPS_INIT;
static const u8 kPredIdx[10][16] = ZW_PRED_IDX_INIT;
extern "C" {
void hc_fdct(i32* b) { fdct4x4(b); }
void hc_idct(i32* b) { idct4x4(b); }
void hc_wht(i32* b) { wht4x4(b); }
void hc_iwht(i32* b) { iwht4x4(b); }
int hc_ttransform(const i32* px) { return t_transform16(px, host::kWeightY); }
void hc_predict4_lut(const u8* e, int mode, u8* out) { for (int k = 0; k < 16; k++) out[k] = (u8)predict4_pixel_lut(e, mode, k, kDTaps, kPredIdx); }
void hc_predict4(const u8* e, int mode, u8* out) { for (int k = 0; k < 16; k++) out[k] = (u8)predict4_pixel(e, mode, k, kPredTab); }
static Matrix mk(const u16* q, const u32* iq, const u32* bias) { Matrix m; for (int i = 0; i < 2; i++) { m.q[i] = q[i]; m.iq[i] = iq[i]; m.bias[i] = bias[i]; } return m; }
int hc_quantize(int coeff, const u16* q, const u32* iq, const u32* bias, int pos) { return quantize_coeff(coeff, mk(q, iq, bias), pos); }
u32 hc_residual_cost(const i32* lv, int ctype, int first, int ctx0, const u8* probs, const u16* lcost) {
  CostCtx cc; cc.probs = probs; cc.level_cost = lcost; return residual_cost(lv, ctype, first, ctx0, cc);
}
void hc_level_costs(const u8* probs, u16* out) {
  for (int i = 0; i < 6528; i++) { int v = i % 68, row = i / 68, ctx = row % 3; out[i] = level_cost_entry(probs + row * 11, ctx, v); }
}
int hc_trellis(i32* coeffs, i32* out, const u16* q, const u32* iq, const u32* bias, const u16* sharpen, u32 lambda, int first,
               const u8* probs, const u16* lcost, int ctype, int ctx0) {
  CostCtx cc; cc.probs = probs; cc.level_cost = lcost;
  return trellis_quantize(coeffs, out, mk(q, iq, bias), sharpen, lambda, first, cc, ctype, ctx0) ? 1 : 0;
}
void hc_token_events(const i16* zz, int t, int first, int ctx, u32* stats /*1056 packed like ProbaStats (no halving)*/) {
  token_events(zz, t, first, ctx, [&](int slot, int bit) { stats[slot] += 0x10000u + (u32)bit; });
}
}
