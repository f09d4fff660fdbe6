// zw_hostcheck.cpp -- TEST INFRASTRUCTURE: compiles the per-lane device primitives
// (image_webp_b200/csrc/zw_prims.cuh, zw_cost.cuh) with the HOST compiler so that the CPU test
// suite can compare them with the oracle on random inputs without a GPU.  Not part of the
// product; the shipped library only runs these functions inside CUDA kernels.
#include <cstring>
#include <vector>
#include "../../image_webp_b200/csrc/zw_boolcoder.cuh"
#include "../../image_webp_b200/csrc/zw_cost.cuh"
#include "../../image_webp_b200/csrc/zw_quad.cuh"
#include "../../image_webp_b200/csrc/zw_dec.cuh"
#include "../../image_webp_b200/csrc/zw_lossless.cuh"
using namespace zw;
static const u16 kPredTab[8][16] = ZW_PRED_TABLE_INIT;
static const u16 kDTaps[32] = ZW_DTAPS_INIT;
static const u8 kPredIdx[10][16] = ZW_PRED_IDX_INIT;
struct HostExec {
  template <class F> void run(F&& f) { for (int q = 0; q < zw::QG; q++) f(q); }
};
struct HostExec32 {
  template <class F> void run(F&& f) { for (int l = 0; l < 32; l++) f(l); }
};
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
int hc_trellis_rolled(i32* coeffs, i32* out, const u16* q, const u32* iq, const u32* bias, const u16* sharpen, u32 lambda, int first,
                      const u8* probs, const u16* lcost, int ctype, int ctx0) {
  CostCtx cc; cc.probs = probs; cc.level_cost = lcost;
  return q_trellis(coeffs, out, mk(q, iq, bias), sharpen, lambda, first, cc, ctype, ctx0) ? 1 : 0;
}
void hc_token_events(const i16* zz, int t, int first, int ctx, u32* stats /*1056 packed like ProbaStats (no halving)*/) {
  token_events(zz, t, first, ctx, [&](int slot, int bit) { stats[slot] += 0x10000u + (u32)bit; });
}
// The five steps of the segment-parallel boolean coder (k_bc_cands / k_bc_trans / k_bc_resolve / k_bc_code / k_bc_fix in
// zw_back.cuh), run serially with the SAME lane-local functions the kernels call; `seg` / `warm` stand for BC_SEG / BC_WARM
// so that small streams cross many segment boundaries.  Returns the stream length; stats[0] = segments, [1] = largest
// candidate set, [2] = carries that crossed a segment start, [3] = candidate sets summed.
size_t hc_boolcode_segmented(const u16* tk, size_t n, u32 seg, u32 warm, u8* out, u32 cap, u32* stats) {
  const u32 J = n == 0 ? 1u : (u32)((n + seg - 1) / seg);
  struct Seg { u32 cand[4]; u64 start_bit, tail; u32 carries; u8 state; };
  std::vector<Seg> S(J);
  std::vector<std::vector<u32>> trans(J);
  u32 maxk = 0, sumk = 0, crossed = 0;
  for (u32 j = 0; j < J; j++) {  // k_bc_cands
    Seg& sg = S[j];
    sg.cand[0] = sg.cand[1] = sg.cand[2] = sg.cand[3] = 0;
    if (j == 0) { sg.cand[3] = 0x80000000u; continue; }
    const u16* w = tk + (size_t)j * seg - warm;
    for (u32 s0 = 127; s0 < 255; s0++) {
      u32 st = s0, add, sh;
      for (u32 i = 0; i < warm; i++) st = bc_step(st, w[i], add, sh);
      sg.cand[(st - 127) >> 5] |= 1u << ((st - 127) & 31);
    }
  }
  for (u32 j = 0; j < J; j++) {  // k_bc_trans
    const Seg& sg = S[j];
    const u32 k = bc_popc(sg.cand[0]) + bc_popc(sg.cand[1]) + bc_popc(sg.cand[2]) + bc_popc(sg.cand[3]);
    maxk = k > maxk ? k : maxk; sumk += k;
    const size_t first = (size_t)j * seg;
    const u32 cnt = n > first ? (u32)((n - first) < seg ? (n - first) : seg) : 0u;
    for (u32 c = 0; c < k; c++) {
      u32 st = 127u + bc_nth_bit(sg.cand[0], sg.cand[1], sg.cand[2], sg.cand[3], c), T = 0, add, sh;
      for (u32 i = 0; i < cnt; i++) { st = bc_step(st, tk[first + i], add, sh); T += sh; }
      trans[j].push_back(st | (T << 8));
    }
  }
  {  // k_bc_resolve
    u32 st = 254;
    u64 bits = 0;
    for (u32 j = 0; j < J; j++) {
      const u32 e = trans[j][bc_rank(S[j].cand[0], S[j].cand[1], S[j].cand[2], S[j].cand[3], st - 127u)];
      S[j].state = (u8)st; S[j].start_bit = bits;
      st = e & 255u; bits += e >> 8;
    }
  }
  size_t bytes = 0;
  for (u32 jj = 0; jj < J; jj++) {  // k_bc_code, in an order that is NOT the stream order
    const u32 j = (jj % 2 == 0) ? (J - 1 - jj / 2) : (jj / 2);
    const size_t first = (size_t)j * seg;
    const u32 cnt = n > first ? (u32)((n - first) < seg ? (n - first) : seg) : 0u;
    BcCoder cd;
    cd.begin(S[j].state, S[j].start_bit, out, cap);
    for (u32 i = 0; i < cnt; i++) cd.put(tk[first + i]);
    if (j + 1 < J) S[j].tail = cd.tail(); else { bytes = cd.flush(); S[j].tail = 0; }
    S[j].carries = cd.carries;
    crossed += cd.carries;
    if (cd.overflow) return (size_t)-1;
  }
  for (u32 j = 1; j < J; j++) bc_fix_boundary(out, S[j].start_bit, S[j - 1].tail, S[j].carries);  // k_bc_fix
  if (stats) { stats[0] = J; stats[1] = maxk; stats[2] = crossed; stats[3] = sumk; }
  return bytes;
}
// The quad (four lanes per macroblock row) luma path of zw_quad.cuh, run over a whole image on the CPU: the lanes of a quad
// are called one after the other at every sync point.  y: padded luma plane (stride 16 * mbw); segmap / seg_qidx: the image's
// segments (segmap == NULL: all base_qidx); pass 1: default probabilities, zero level costs, no trellis; pass 2: the
// image's probabilities / level costs and, for method >= 4, trellis.  uvnz: pass-2 chroma has_coeffs bits per macroblock.
// out: one MbRecord per macroblock with the luma-owned fields filled (ymode, bmodes, levels[0..16]; pass 2 also skip,
// top_nz, left_nz).
void hc_quad_luma_image(const u8* y, int mbw, int mbh, int pass, int method, int base_qidx, const u8* segmap, const u8* seg_qidx,
                        const u8* probs, const u16* lcost, const u8* uvnz, MbRecord* out) {
  static u8 pidx[10][16];
  static bool init = false;
  if (!init) {
    memcpy(pidx, kPredIdx, sizeof(pidx));
    for (int i = 0; i < 16; i++) pidx[1][i] = (u8)(32 + i);  // TM pixels live in dtab[32 + n]
    init = true;
  }
  QuadConst K; K.pidx = pidx; K.dtaps = kDTaps;
  std::vector<MbBottom> bottom((size_t)mbw * mbh);
  std::vector<u16> nz_after((size_t)mbw * mbh);
  const int pw = mbw * 16;
  HostExec X;
  for (int mby = 0; mby < mbh; mby++) {
    QuadScratch S;
    memset(&S, 0, sizeof(S));
    u32 left_nz = 0;
    for (int mbx = 0; mbx < mbw; mbx++) {
      const int mb = mby * mbw + mbx;
      const SegParams SP = make_segparams(segmap ? seg_qidx[segmap[mb]] : base_qidx);
      // load_luma_mb (zw_search.cuh): source, top / top-right / left borders
      for (int r = 0; r < 16; r++) memcpy(&S.src_y[r * 16], y + (size_t)(mby * 16 + r) * pw + mbx * 16, 16);
      if (mby == 0) { for (int k = 0; k < 32; k++) S.yws[k] = 127; }
      else {
        const MbBottom* bt = &bottom[mb - mbw];
        for (int k = 0; k < 16; k++) S.yws[1 + k] = bt->y[k];
        for (int k = 16; k < 20; k++) S.yws[1 + k] = (mbx == mbw - 1) ? bt->y[15] : (bt + 1)->y[k - 16];
      }
      for (int k = 0; k < 16; k++) S.yws[(1 + k) * 32] = (mbx == 0) ? 129 : S.left_y[1 + k];
      S.yws[0] = (mby == 0) ? 127 : (mbx == 0 ? 129 : S.left_y[0]);
      for (int k = 0; k < 12; k++) { const int r = 4 * (1 + k / 4), c = 17 + (k & 3); S.yws[r * 32 + c] = S.yws[c]; }
      memset(S.lv, 0, sizeof(S.lv));
      QuadMbIn in;
      in.SP = &SP;
      in.cc.probs = pass == 1 ? host::kCoeffProbs : probs;
      in.cc.level_cost = pass == 1 ? nullptr : lcost;
      in.i4_modes = method <= 1 ? 0 : (method <= 3 ? 3 : (method == 4 ? 4 : 10));
      in.i4_always = method >= 5;
      in.trellis = pass == 2 && method >= 4;
      in.mbx = mbx; in.mby = mby;
      in.in_top_nz = (pass == 2 && mby > 0) ? nz_after[mb - mbw] : 0;
      in.in_left_nz = left_nz;
      const QuadLumaOut L = quad_luma_mb<QG>(X, S, K, in);
      MbRecord& r = out[mb];
      memset(&r, 0, sizeof(r));
      r.ymode = L.use_i4 ? 4 : (u8)L.mode16;
      if (L.use_i4) memcpy(r.bmodes, S.bmodes, 16);
      r.segment = segmap ? segmap[mb] : 0;
      bool skip = false;
      u32 out_top = 0, out_left = 0;
      if (pass == 2) {
        skip = !(L.simple_nz || uvnz[mb] != 0);
        q_complexity_after(L.use_i4, skip, L.y2nz, L.ynz, uvnz[mb], in.in_top_nz, left_nz, out_top, out_left);
        r.skip = skip; r.top_nz = (u16)in.in_top_nz; r.left_nz = (u16)left_nz;
        nz_after[mb] = (u16)out_top;
      } else {
        r.top_nz = (u16)L.ynz; r.left_nz = (u16)((L.y2nz ? 1 : 0) | (L.simple_nz ? 2 : 0));  // parked for k_finish1
      }
      if (!skip) memcpy(r.levels, S.lv, sizeof(S.lv));
      left_nz = out_left;
      for (int k = 0; k < 17; k++) S.left_y[k] = S.yws[k * 32 + 16];
      for (int k = 0; k < 16; k++) bottom[mb].y[k] = S.yws[16 * 32 + 1 + k];
    }
  }
}

// The on-device decoder (zw_dec.cuh), the SAME source the kernels run, lane by lane on the host: one VP8 frame ->
// filtered planes (Y | U | V padded), macroblock info words, RGB (fancy or simple) and the squared error against `src`.
// Returns the decoder status.
int hc_decode(const u8* data, size_t len, u32 width, u32 height, int fancy, u8* planes, u32* mbinfo, u8* rgb, const u8* src, u32 src_bpp,
              u32* st_out /*[12]*/, u64* sse) {
  DecImage D;
  memset(&D, 0, sizeof(D));
  D.data_off = 0; D.data_len = (u32)len; D.width = width; D.height = height; D.mbw = (width + 15) / 16; D.mbh = (height + 15) / 16;
  D.src_bpp = src_bpp;
  std::vector<u16> topnz(1024);
  std::vector<u32> topmodes(1024);
  std::vector<uint4> recbuf(((size_t)D.mbw * D.mbh * sizeof(MbRecord)) / 16 + 1);
  DecState st;
  memset(&st, 0, sizeof(st));
  DecParams P;
  memset(&P, 0, sizeof(P));
  P.img = &D; P.st = &st; P.n_img = 1; P.fancy = fancy; P.bytes = data; P.planes = planes; P.mbinfo = mbinfo;
  P.rec = reinterpret_cast<MbRecord*>(recbuf.data()); P.rgb = rgb; P.src = src;
  static DecParseShared SP;
  static DecReconShared SR;
  HostExec32 X;
  dec_parse_frame(X, SP, P, D, 0, len, st, topnz.data(), topmodes.data());
  if (st.status == 0) {
    for (u32 y = 0; y < D.mbh; y++)
      for (u32 xm = 0; xm < D.mbw; xm++) dec_recon_mb(X, SR, P, D, st, (int)xm, (int)y, kPredTab);
    if (st.filter_level != 0)
      for (u32 y = 0; y < D.mbh; y++)
        for (u32 xm = 0; xm < D.mbw; xm++) dec_filter_mb(X, st, D, planes, (int)xm, (int)y, mbinfo[((size_t)y * D.mbw + xm) * 4]);
  }
  u64 t = 0;
  if (st.status == 0)
    for (u32 row = 0; row < height; row++)
      for (u32 xx = 0; xx < width; xx++) {
        u32 px[3];
        dec_rgb_pixel(planes, D, fancy, row, xx, px);
        if (rgb) for (int k = 0; k < 3; k++) rgb[((size_t)row * width + xx) * 3 + k] = (u8)px[k];
        if (src && src_bpp) t += dec_pixel_sse(src + ((size_t)row * width + xx) * src_bpp, src_bpp, px);
      }
  if (sse) *sse = t;
  if (st_out) memcpy(st_out, &st, 12 * sizeof(u32));
  return (int)st.status;
}

// ---- lossless (zw_lossless.cuh): the per-pixel functions, the Huffman construction and the header writer the kernels call,
// driven serially.  The tile scans, the shared-memory bit packing and the output layout are the kernels' own and are only
// checked on a GPU.  Returns the stream length in bytes; hist_out [4][280], codes_out [4][280] (length << 16 | code).
size_t hc_lossless(const u8* src, u32 width, u32 height, u32 bpp, u32 color, u32 flags, u8* out, size_t cap, u32* hist_out, u32* codes_out,
                   u32* res_out, u16* desc_out) {
  const u32 npx = width * height;
  std::vector<u32> res(npx);
  std::vector<u16> desc(npx);
  for (u32 i = 0; i < npx; i++) res[i] = ll_residual(src, i, i % width, i / width, width, bpp, color, flags);
  u32 head = 0;
  std::vector<u32> hist(4 * 280, 0);
  for (u32 i = 0; i < npx; i++) {
    if (i == 0 || res[i] != res[i - 1]) head = i;
    const u32 d = ll_token_desc(i, head, i + 1 == npx || res[i + 1] != res[i]);
    desc[i] = (u16)d;
    if (d & 1) {
      hist[280 + ((res[i] >> 8) & 255)]++;
      if (ll_is_color(color)) { hist[res[i] & 255]++; hist[560 + ((res[i] >> 16) & 255)]++; }
      if (ll_is_alpha(color)) hist[840 + (res[i] >> 24)]++;
    }
    if (d >> 1) { u32 sym, eb, ev; ll_run_symbol(d >> 1, sym, eb, ev); hist[280 + sym]++; }
  }
  std::vector<u32> words(cap / 4 + 16, 0);
  LlBits w; w.w = words.data(); w.pos = 0;
  ll_write_prefix(w, width, height, color, flags);
  static LlHuffScratch S;
  u8 lengths[4][280]; u16 codes[4][280];
  memset(lengths, 0, sizeof(lengths)); memset(codes, 0, sizeof(codes));
  const int order[4] = {1, 0, 2, 3};
  for (int k = 0; k < 4; k++) {
    const int c = order[k];
    const bool built = c == 1 || (ll_is_color(color) && (c == 0 || c == 2)) || (ll_is_alpha(color) && c == 3);
    if (built) ll_write_huffman_tree(w, &hist[c * 280], c == 1 ? 280 : 256, lengths[c], codes[c], S);
    else if (c == 3) ll_write_single_entry_tree(w, (flags & LL_FLAG_PREDICTOR) ? 0u : 255u);
    else ll_write_single_entry_tree(w, 0);
  }
  ll_write_single_entry_tree(w, 1);
  std::vector<u32> tab(4 * 280);
  for (int c = 0; c < 4; c++) for (int k = 0; k < 280; k++) tab[c * 280 + k] = ((u32)lengths[c][k] << 16) | codes[c][k];
  auto table = [&](u32 ch, u32 sym) { return tab[ch * 280 + sym]; };
  u64 total = w.pos;
  for (u32 i = 0; i < npx; i++) total += ll_pixel_bits(res[i], desc[i], color, table);
  if ((total + 7) / 8 > cap) return (size_t)((total + 7) / 8);
  for (u32 i = 0; i < npx; i++) {
    const u32 px = res[i], d = desc[i];
    if (d & 1) {
      u32 e = table(1, (px >> 8) & 255); w.put(e & 0xFFFF, e >> 16);
      if (ll_is_color(color)) { e = table(0, px & 255); w.put(e & 0xFFFF, e >> 16); e = table(2, (px >> 16) & 255); w.put(e & 0xFFFF, e >> 16); }
      if (ll_is_alpha(color)) { e = table(3, px >> 24); w.put(e & 0xFFFF, e >> 16); }
    }
    if (d >> 1) {
      u32 sym, eb, ev; ll_run_symbol(d >> 1, sym, eb, ev);
      const u32 e = table(1, sym);
      w.put(e & 0xFFFF, e >> 16); w.put(ev, eb);
    }
  }
  const size_t bytes = (size_t)((w.pos + 7) / 8);
  memcpy(out, words.data(), bytes);
  if (hist_out) memcpy(hist_out, hist.data(), 4 * 280 * 4);
  if (codes_out) memcpy(codes_out, tab.data(), 4 * 280 * 4);
  if (res_out) memcpy(res_out, res.data(), (size_t)npx * 4);
  if (desc_out) memcpy(desc_out, desc.data(), (size_t)npx * 2);
  return bytes;
}
int hc_ll_huffman(const u32* freq, u32 n, u32 limit, u8* lengths, u16* codes) {
  static LlHuffScratch S;
  return ll_build_huffman(freq, n, lengths, codes, limit, S) ? 1 : 0;
}
}
