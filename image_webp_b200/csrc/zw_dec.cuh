// zw_dec.cuh -- on-device VP8 key-frame DECODER, used as the batch verifier: every bitstream of a
// batch is parsed, reconstructed, loop-filtered, converted to RGB and scored against its source on the GPU.
//
// Reference (file:line under /root/reference):
//   boolean decoder        src/decoder/bit_reader.rs:29-250, :254-438         (rd_load / rd_bit)
//   frame header           src/decoder/vp8.rs:553-679 (+ :387, :421, :452, :506, :520)   (dec_parse_header)
//   macroblock header      src/decoder/vp8.rs:681-734                         (dec_parse_mb)
//   coefficients           src/decoder/vp8.rs:872-1058, residual data :1060-1170        (dec_read_coeffs, dec_parse_mb)
//   prediction + residue   src/decoder/vp8.rs:736-870; src/common/prediction.rs          (dec_reconstruct_mb)
//   loop filter            src/decoder/vp8.rs:1172-1348, :1470-1524; src/decoder/loop_filter.rs      (dec_filter_mb)
//   YUV -> RGB             src/decoder/yuv.rs:36-78, :82-399 (bilinear), :402-530 (nearest)          (k_dec_rgb)
//
// Mapping.  The two bitstream partitions of an image are strictly serial (every symbol's interval depends on the one
// before), so ONE WARP owns ONE IMAGE: lane 0 parses a macroblock (header from the first partition, coefficients from
// the token partition of its row) into shared memory, then the 32 lanes reconstruct it -- 16 luma + 8 chroma 4x4 blocks,
// one block per lane (inverse transforms, whole-block predictors, the sub-block wavefront x + 2y for B_PRED) -- and
// store it to the plane arena.  A second sweep of the same warp runs the loop filter in macroblock raster order, one
// lane per pixel row / column of an edge (16 luma + 8 + 8 chroma lanes).  The batch supplies the parallelism: 1024
// images = 1024 warps.  A separate data-parallel kernel does the chroma upsampling, the colour conversion and the
// squared error against the source pixels.
//
// The per-image work is written against an executor X (X::run(f) = "every lane calls f(lane), then the warp
// synchronises") and compiles for the host too: tests/hostcheck runs the same source lane by lane on the CPU and
// compares planes, modes and RGB with the decoder oracle before the code reaches a GPU.
#ifndef ZW_DEC_CUH
#define ZW_DEC_CUH
#include "zw_types.cuh"

namespace zw {

enum { ZWD_OK = 0, ZWD_BITSTREAM = 1, ZWD_UNSUPPORTED = 2, ZWD_MAGIC = 3, ZWD_COLORSPACE = 4, ZWD_TRUNCATED = 5, ZWD_CONTAINER = 6,
       ZWD_DIMENSIONS = 7 };

// Host-filled description of one file of a decode batch.
struct DecImage {
  u64 data_off;     // the VP8 frame (container already stripped) in the byte arena
  u32 data_len;
  u32 width, height, mbw, mbh;  // as the host read them from the frame header (or knows them from the source)
  u32 mb_off;       // first macroblock in mbinfo
  u32 col_off;      // first macroblock column in the top-context scratch
  u32 src_bpp;      // verify: bytes per source pixel (1 L8, 2 La8, 3 Rgb8, 4 Rgba8), 0 = no source
  u64 plane_off;    // padded Y | U | V planes in the plane arena
  u64 rgb_off;      // decoded RGB (width * height * 3) in the RGB arena
  u64 src_off;      // verify: source pixels in the source arena
};

// Device-written per-image result.
struct DecState {
  u32 status;
  u32 filter_type, filter_level, sharpness, num_partitions, segments_enabled, update_map, lf_adj, has_skip_prob, prob_skip_false,
      version, pixel_type;
  u64 sse_rgb;      // sum over the width x height x 3 colour samples of (decoded - source)^2
};

struct DecParams {
  const DecImage* img;
  DecState* st;
  u32 n_img;
  int fancy;            // 1: bilinear chroma upsampling (the reference's default), 0: nearest
  const u8* bytes;      // file arena
  const u64* file_off;  // optional (verify after encode): device-side offsets of the files in `bytes` ...
  const ImageState* enc_st;  // ... and their sizes; data_off / data_len are then taken from here (+ 20-byte RIFF wrap)
  u8* planes;
  u32* mbinfo;          // [n_mb][4]: flags (luma_mode | chroma_mode << 3 | segment << 5 | skipped << 7 | non_zero_dct << 8), bpred nibbles x2, 0
  u16* topnz;           // [columns] complexity left behind by the row above (bit 0 y2, 1..4 y, 5..6 u, 7..8 v)
  u32* topmodes;        // [columns] its bottom four sub-block modes
  u8* rgb;              // may be null (verify only)
  const u8* src;        // may be null (decode only)
};

// ---- boolean decoder: libwebp's VP8GetBitAlt with 56-bit refills, range - 1 stored (bit_reader.rs:29-147) ----
struct DecReader {
  u64 value, pos, end;
  u32 range;
  i32 bits;
  u32 eof, pad;
};
ZW_HD int dec_clz(u32 v) {
#if defined(__CUDA_ARCH__)
  return __clz((int)v);
#else
  return __builtin_clz(v);
#endif
}
ZW_HD void rd_load(DecReader& r, const u8* bytes) {  // load_new_bytes :75 / load_final_bytes :59
  if (r.end - r.pos >= 7) {
    u64 in = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) in = (in << 8) | (u64)bytes[r.pos + i];
    r.value = in | (r.value << 56);
    r.bits += 56;
    r.pos += 7;
  } else if (r.pos < r.end) {
    r.bits += 8;
    r.value = (u64)bytes[r.pos] | (r.value << 8);
    r.pos++;
  } else if (!r.eof) {
    r.value <<= 8;
    r.bits += 8;
    r.eof = 1;
  } else r.bits = 0;
}
ZW_HD void rd_init(DecReader& r, const u8* bytes, u64 off, u64 len) {
  r.value = 0; r.range = 255 - 1; r.bits = -8; r.pos = off; r.end = off + len; r.eof = 0; r.pad = 0;
  rd_load(r, bytes);
}
ZW_HD int rd_bit(DecReader& r, const u8* bytes, u32 prob) {  // get_bit :120
  u32 range = r.range;
  if (r.bits < 0) rd_load(r, bytes);
  const int p = r.bits;
  const u32 split = (range * prob) >> 8;
  const u32 v = (u32)(r.value >> p);
  const int bit = v > split;
  if (bit) { range -= split; r.value -= ((u64)split + 1) << p; }
  else range = split + 1;
  const int shift = 7 ^ (31 ^ dec_clz(range));
  range <<= shift;
  r.bits -= shift;
  r.range = range - 1;
  return bit;
}
ZW_HD u32 rd_literal(DecReader& r, const u8* bytes, int n) {  // read_literal :385 (u8 arithmetic: n <= 8)
  u32 v = 0;
  for (int i = 0; i < n; i++) v = ((v << 1) | (u32)rd_bit(r, bytes, 128)) & 255u;
  return v;
}
ZW_HD i32 rd_optional_signed(DecReader& r, const u8* bytes, int n) {  // :395
  if (!rd_bit(r, bytes, 128)) return 0;
  const i32 m = (i32)rd_literal(r, bytes, n);
  return rd_bit(r, bytes, 128) ? -m : m;
}
// read_with_tree (:409) over the i8 trees of src/common/types.rs (a leaf is stored as -value; 0 is the leaf "value 0")
ZW_HD int rd_tree(DecReader& r, const u8* bytes, const i8* tree, const u8* probs) {
  int i = 0;
  for (;;) {
    const int nx = tree[i + rd_bit(r, bytes, probs[i >> 1])];
    if (nx <= 0) return -nx;
    i = nx;
  }
}

// What the lanes of the warp that owns an image share (one per warp in shared memory; a plain struct on the host).
struct DecShared {
  u8 probs[4][8][3][11];  // token probabilities of the frame (types.rs:338, updated :387)
  DecReader rd[9];        // [0] first partition, [1 + p] token partition p
  i32 coef[24][16];       // dequantised coefficients, then residuals: 0..15 Y, 16..19 U, 20..23 V (natural order)
  i32 y2[16];
  u8 nflag[24];           // block has coefficients past its first position (full inverse DCT)
  u8 yws[17 * 32];        // bordered work buffers (prediction.rs LUMA_STRIDE / CHROMA_STRIDE = 32)
  u8 uws[9 * 32];
  u8 vws[9 * 32];
  i16 quant[4][6];        // per segment: ydc, yac, y2dc, y2ac, uvdc, uvac (:452-504)
  i8 seg_quant[4], seg_lf[4];
  u8 seg_delta;           // segment values are deltas
  u8 seg_probs[3];
  i32 ref_delta0, mode_delta0;
  u32 width, height, mbw, mbh;
  u8 segments_enabled, update_map, filter_type, filter_level, sharpness, lf_adj, has_skip_prob, prob_skip_false, num_partitions;
  // the macroblock in flight
  u8 bpred[16];
  u8 left_bpred[4];
  u8 luma_mode, chroma_mode, segment, skipped, nzdct;
  u32 left_nz;
  u32 status;
};

// the trees of src/common/types.rs:191-205, :332 (a leaf is stored as -value)
#define ZW_DT_SEG {2, 4, 0, -1, -2, -3}
#define ZW_DT_YMODE {-4, 2, 4, 6, 0, -1, -2, -3}
#define ZW_DT_BMODE {0, 2, -1, 4, -2, 6, 8, 12, -3, 10, -5, -6, -4, 14, -7, 16, -8, -9}
#define ZW_DT_UV {0, 2, -1, 4, -2, -3}

// read_frame_header (vp8.rs:553-679), lane 0.  Leaves the readers, probabilities, quantisers and filter settings in S.
ZW_HD int dec_parse_header(DecShared& S, const u8* bytes, u64 off, u64 len, const DecImage& D, DecState& st) {
  if (len < 3) return ZWD_TRUNCATED;
  const u32 tag = (u32)bytes[off] | ((u32)bytes[off + 1] << 8) | ((u32)bytes[off + 2] << 16);
  if (tag & 1) return ZWD_UNSUPPORTED;
  st.version = (tag >> 1) & 7;
  const u64 first_size = tag >> 5;
  if (len < 6) return ZWD_TRUNCATED;
  if (!(bytes[off + 3] == 0x9d && bytes[off + 4] == 0x01 && bytes[off + 5] == 0x2a)) return ZWD_MAGIC;
  if (len < 10) return ZWD_TRUNCATED;
  const u32 w = ((u32)bytes[off + 6] | ((u32)bytes[off + 7] << 8)) & 0x3FFF, h = ((u32)bytes[off + 8] | ((u32)bytes[off + 9] << 8)) & 0x3FFF;
  if (w != D.width || h != D.height) return ZWD_DIMENSIONS;  // the arenas were laid out for D's dimensions
  S.width = w; S.height = h; S.mbw = (w + 15) / 16; S.mbh = (h + 15) / 16;
  u64 rpos = 10;
  if (len - rpos < first_size || first_size == 0) return ZWD_TRUNCATED;
  DecReader b;
  rd_init(b, bytes, off + rpos, first_size);
  rpos += first_size;
  const u32 color_space = rd_literal(b, bytes, 1);
  st.pixel_type = rd_literal(b, bytes, 1);
  if (color_space != 0) return ZWD_COLORSPACE;
  S.segments_enabled = (u8)rd_bit(b, bytes, 128);
  S.update_map = 0; S.seg_delta = 0;
  for (int i = 0; i < 4; i++) { S.seg_quant[i] = 0; S.seg_lf[i] = 0; }
  for (int i = 0; i < 3; i++) S.seg_probs[i] = 255;
  if (S.segments_enabled) {  // read_segment_updates :520
    S.update_map = (u8)rd_bit(b, bytes, 128);
    if (rd_bit(b, bytes, 128)) {
      S.seg_delta = (u8)!rd_bit(b, bytes, 128);
      for (int i = 0; i < 4; i++) S.seg_quant[i] = (i8)rd_optional_signed(b, bytes, 7);
      for (int i = 0; i < 4; i++) S.seg_lf[i] = (i8)rd_optional_signed(b, bytes, 6);
    }
    if (S.update_map)
      for (int i = 0; i < 3; i++) S.seg_probs[i] = rd_bit(b, bytes, 128) ? (u8)rd_literal(b, bytes, 8) : (u8)255;
    if (b.eof) return ZWD_BITSTREAM;
  }
  S.filter_type = (u8)rd_bit(b, bytes, 128);
  S.filter_level = (u8)rd_literal(b, bytes, 6);
  S.sharpness = (u8)rd_literal(b, bytes, 3);
  S.lf_adj = (u8)rd_bit(b, bytes, 128);
  S.ref_delta0 = 0; S.mode_delta0 = 0;
  if (S.lf_adj) {  // read_loop_filter_adjustments :506 (only ref_delta[0] / mode_delta[0] matter for key frames)
    if (rd_bit(b, bytes, 128)) {
      for (int i = 0; i < 4; i++) { const i32 v = rd_optional_signed(b, bytes, 6); if (i == 0) S.ref_delta0 = v; }
      for (int i = 0; i < 4; i++) { const i32 v = rd_optional_signed(b, bytes, 6); if (i == 0) S.mode_delta0 = v; }
    }
    if (b.eof) return ZWD_BITSTREAM;
  }
  const int nparts = 1 << rd_literal(b, bytes, 2);
  S.num_partitions = (u8)nparts;
  if (b.eof) return ZWD_BITSTREAM;
  {  // init_partitions :421
    u64 sizes[8];
    if (nparts > 1) {
      if (len - rpos < (u64)(3 * nparts - 3)) return ZWD_TRUNCATED;
      for (int i = 0; i < nparts - 1; i++) {
        sizes[i] = (u64)bytes[off + rpos] | ((u64)bytes[off + rpos + 1] << 8) | ((u64)bytes[off + rpos + 2] << 16);
        rpos += 3;
      }
    }
    for (int i = 0; i < nparts - 1; i++) {
      if (len - rpos < sizes[i]) return ZWD_TRUNCATED;
      rd_init(S.rd[1 + i], bytes, off + rpos, sizes[i]);
      rpos += sizes[i];
    }
    rd_init(S.rd[nparts], bytes, off + rpos, len - rpos);
  }
  {  // read_quantization_indices :452
    const i32 yac_abs = (i32)rd_literal(b, bytes, 7);
    const i32 ydc_d = rd_optional_signed(b, bytes, 4), y2dc_d = rd_optional_signed(b, bytes, 4), y2ac_d = rd_optional_signed(b, bytes, 4);
    const i32 uvdc_d = rd_optional_signed(b, bytes, 4), uvac_d = rd_optional_signed(b, bytes, 4);
    const int n = S.segments_enabled ? 4 : 1;
    for (int i = 0; i < n; i++) {
      const i32 base = S.segments_enabled ? (S.seg_delta ? (i32)S.seg_quant[i] + yac_abs : (i32)S.seg_quant[i]) : yac_abs;
      i32 v;
      S.quant[i][0] = ZW_TAB(kDcQuant)[imin(imax(base + ydc_d, 0), 127)];
      S.quant[i][1] = ZW_TAB(kAcQuant)[imin(imax(base, 0), 127)];
      S.quant[i][2] = (i16)(ZW_TAB(kDcQuant)[imin(imax(base + y2dc_d, 0), 127)] * 2);
      v = (i32)ZW_TAB(kAcQuant)[imin(imax(base + y2ac_d, 0), 127)] * 155 / 100;
      S.quant[i][3] = (i16)(v < 8 ? 8 : v);
      v = ZW_TAB(kDcQuant)[imin(imax(base + uvdc_d, 0), 127)];
      S.quant[i][4] = (i16)(v > 132 ? 132 : v);
      S.quant[i][5] = ZW_TAB(kAcQuant)[imin(imax(base + uvac_d, 0), 127)];
    }
    if (b.eof) return ZWD_BITSTREAM;
  }
  (void)rd_literal(b, bytes, 1);  // refresh entropy probs
  for (int i = 0; i < 1056; i++) {  // update_token_probabilities :387 (defaults were copied into S.probs by all lanes)
    if (rd_bit(b, bytes, ZW_TAB(kCoeffUpdateProbs)[i])) (&S.probs[0][0][0][0])[i] = (u8)rd_literal(b, bytes, 8);
  }
  if (b.eof) return ZWD_BITSTREAM;
  S.has_skip_prob = (u8)(rd_literal(b, bytes, 1) == 1);
  S.prob_skip_false = S.has_skip_prob ? (u8)rd_literal(b, bytes, 8) : (u8)0;
  if (b.eof) return ZWD_BITSTREAM;
  S.rd[0] = b;
  st.filter_type = S.filter_type; st.filter_level = S.filter_level; st.sharpness = S.sharpness; st.num_partitions = S.num_partitions;
  st.segments_enabled = S.segments_enabled; st.update_map = S.update_map; st.lf_adj = S.lf_adj; st.has_skip_prob = S.has_skip_prob;
  st.prob_skip_false = S.prob_skip_false;
  return ZWD_OK;
}

// read_coefficients (vp8.rs:872-1058).  Returns 1 / 0 (coefficients past `first` or not), -1 on a bitstream error.
ZW_HD int dec_read_coeffs(DecReader& r, const u8* bytes, const u8 (*probs)[3][11], int first, int ctx, i32 dcq, i32 acq, i32* block) {
  int n = first;
  const u8* prob = probs[ZW_TAB(kCoeffBands)[n]][ctx];
  while (n < 16) {
    if (!rd_bit(r, bytes, prob[0])) break;
    while (!rd_bit(r, bytes, prob[1])) {
      n++;
      if (n >= 16) return r.eof ? -1 : 1;
      prob = probs[ZW_TAB(kCoeffBands)[n]][0];
    }
    i32 v;
    int next_ctx;
    if (!rd_bit(r, bytes, prob[2])) { v = 1; next_ctx = 1; }
    else {
      if (!rd_bit(r, bytes, prob[3])) {
        if (!rd_bit(r, bytes, prob[4])) v = 2;
        else v = 3 + rd_bit(r, bytes, prob[5]);
      } else if (!rd_bit(r, bytes, prob[6])) {
        if (!rd_bit(r, bytes, prob[7])) v = 5 + rd_bit(r, bytes, 159);
        else { v = 7 + 2 * rd_bit(r, bytes, 165); v += rd_bit(r, bytes, 145); }
      } else {
        const int bit1 = rd_bit(r, bytes, prob[8]);
        const int bit0 = rd_bit(r, bytes, prob[9 + bit1]);
        const int cat = 2 * bit1 + bit0;
        i32 extra = 0;
        for (int k = 0; k < 12; k++) {
          const u32 cp = ZW_TAB(kProbDctCat)[(2 + cat) * 12 + k];
          if (cp == 0) break;
          extra = extra + extra + rd_bit(r, bytes, cp);
        }
        v = 3 + (8 << cat) + extra;
      }
      next_ctx = 2;
    }
    if (rd_bit(r, bytes, 128)) v = -v;
    const int zz = ZW_TAB(kZigzag)[n];
    block[zz] = v * (zz > 0 ? acq : dcq);
    n++;
    if (n < 16) prob = probs[ZW_TAB(kCoeffBands)[n]][next_ctx];
  }
  if (r.eof) return -1;
  return n > first;
}

// One macroblock: read_macroblock_header (:681) + read_residual_data (:1060) or the skip bookkeeping (:1546-1556).
// Lane 0.  S.coef / S.y2 / S.nflag are zero on entry.
ZW_HD int dec_parse_mb(DecShared& S, const u8* bytes, const DecParams& P, const DecImage& D, int mbx, int part) {
  const i8 T_SEG[6] = ZW_DT_SEG, T_YMODE[8] = ZW_DT_YMODE, T_BMODE[18] = ZW_DT_BMODE, T_UV[6] = ZW_DT_UV;
  DecReader b = S.rd[0];
  u32 topm = P.topmodes[D.col_off + mbx];
  S.segment = (S.segments_enabled && S.update_map) ? (u8)rd_tree(b, bytes, T_SEG, S.seg_probs) : (u8)0;
  S.skipped = S.has_skip_prob ? (u8)rd_bit(b, bytes, S.prob_skip_false) : (u8)0;
  S.luma_mode = (u8)rd_tree(b, bytes, T_YMODE, ZW_TAB(kKfYmodeProbs));
  if (S.luma_mode == 4) {
    for (int y = 0; y < 4; y++)
      for (int x = 0; x < 4; x++) {
        const int t = (topm >> (8 * x)) & 255, l = S.left_bpred[y];
        const u8 bm = (u8)rd_tree(b, bytes, T_BMODE, &ZW_TAB(kKfBmodeProbs)[(t * 10 + l) * 9]);
        S.bpred[x + y * 4] = bm;
        topm = (topm & ~(255u << (8 * x))) | ((u32)bm << (8 * x));
        S.left_bpred[y] = bm;
      }
  } else {
    const u8 m = S.luma_mode == 0 ? 0 : (S.luma_mode == 1 ? 2 : (S.luma_mode == 2 ? 3 : 1));  // into_intra: DC, VE, HE, TM
    for (int i = 0; i < 12; i++) S.bpred[i] = 0;
    for (int i = 0; i < 4; i++) { S.bpred[12 + i] = m; S.left_bpred[i] = m; }
    topm = m * 0x01010101u;
  }
  S.chroma_mode = (u8)rd_tree(b, bytes, T_UV, ZW_TAB(kKfUvModeProbs));
  P.topmodes[D.col_off + mbx] = topm;
  S.rd[0] = b;
  if (b.eof) return ZWD_BITSTREAM;
  u32 tnz = P.topnz[D.col_off + mbx], lnz = S.left_nz;
  S.nzdct = 0;
  if (S.skipped) {
    if (S.luma_mode != 4) { tnz &= ~1u; lnz &= ~1u; }
    tnz &= 1u; lnz &= 1u;
  } else {
    DecReader r = S.rd[1 + part];
    const i16* q = S.quant[S.segment];
    int plane = S.luma_mode == 4 ? 3 : 1;  // Plane::YCoeff0 / Y2 (types.rs:48-57: YCoeff1 0, Y2 1, Chroma 2, YCoeff0 3)
    if (plane == 1) {
      const int ctx = (int)(tnz & 1) + (int)(lnz & 1);
      const int n = dec_read_coeffs(r, bytes, S.probs[1], 0, ctx, q[2], q[3], S.y2);
      if (n < 0) return ZWD_BITSTREAM;
      tnz = (tnz & ~1u) | (u32)n; lnz = (lnz & ~1u) | (u32)n;
      plane = 0;
      S.nflag[0] |= 0x80;  // "Y2 present": the lanes run the inverse WHT
    }
    const int first = plane == 0 ? 1 : 0;
    for (int y = 0; y < 4; y++) {
      u32 l = (lnz >> (1 + y)) & 1;
      for (int x = 0; x < 4; x++) {
        const int ctx = (int)((tnz >> (1 + x)) & 1) + (int)l;
        const int n = dec_read_coeffs(r, bytes, S.probs[plane], first, ctx, q[0], q[1], S.coef[x + y * 4]);
        if (n < 0) return ZWD_BITSTREAM;
        S.nflag[x + y * 4] |= (u8)n;
        l = (u32)n;
        tnz = (tnz & ~(2u << x)) | ((u32)n << (1 + x));
      }
      lnz = (lnz & ~(2u << y)) | (l << (1 + y));
    }
    for (int pl = 0; pl < 2; pl++) {
      const int j = 5 + 2 * pl;
      for (int y = 0; y < 2; y++) {
        u32 l = (lnz >> (j + y)) & 1;
        for (int x = 0; x < 2; x++) {
          const int i = x + y * 2 + 16 + 4 * pl;
          const int ctx = (int)((tnz >> (j + x)) & 1) + (int)l;
          const int n = dec_read_coeffs(r, bytes, S.probs[2], 0, ctx, q[4], q[5], S.coef[i]);
          if (n < 0) return ZWD_BITSTREAM;
          S.nflag[i] |= (u8)n;
          l = (u32)n;
          tnz = (tnz & ~(1u << (j + x))) | ((u32)n << (j + x));
        }
        lnz = (lnz & ~(1u << (j + y))) | (l << (j + y));
      }
    }
    S.rd[1 + part] = r;
  }
  P.topnz[D.col_off + mbx] = (u16)tnz;
  S.left_nz = lnz;
  return ZWD_OK;
}

// inverse DCT with the reference's 64-bit products (transform.rs:35-79): a hostile stream can carry coefficients the
// encoder never produces
ZW_HD void dec_idct4x4(i32* b) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const i64 a1 = (i64)b[i] + b[8 + i], b1 = (i64)b[i] - b[8 + i];
    i64 t1 = ((i64)b[4 + i] * 35468) >> 16;
    i64 t2 = (i64)b[12 + i] + (((i64)b[12 + i] * 20091) >> 16);
    const i64 c1 = t1 - t2;
    t1 = (i64)b[4 + i] + (((i64)b[4 + i] * 20091) >> 16);
    t2 = ((i64)b[12 + i] * 35468) >> 16;
    const i64 d1 = t1 + t2;
    b[i] = (i32)(a1 + d1); b[4 + i] = (i32)(b1 + c1); b[12 + i] = (i32)(a1 - d1); b[8 + i] = (i32)(b1 - c1);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const i64 a1 = (i64)b[4 * i] + b[4 * i + 2], b1 = (i64)b[4 * i] - b[4 * i + 2];
    i64 t1 = ((i64)b[4 * i + 1] * 35468) >> 16;
    i64 t2 = (i64)b[4 * i + 3] + (((i64)b[4 * i + 3] * 20091) >> 16);
    const i64 c1 = t1 - t2;
    t1 = (i64)b[4 * i + 1] + (((i64)b[4 * i + 1] * 20091) >> 16);
    t2 = ((i64)b[4 * i + 3] * 35468) >> 16;
    const i64 d1 = t1 + t2;
    b[4 * i] = (i32)((a1 + d1 + 4) >> 3); b[4 * i + 3] = (i32)((a1 - d1 + 4) >> 3);
    b[4 * i + 1] = (i32)((b1 + c1 + 4) >> 3); b[4 * i + 2] = (i32)((b1 - c1 + 4) >> 3);
  }
}

// Whole-block predictors (16x16 luma, 8x8 chroma) for the 4x4 block (bx, by): prediction.rs:164-324
ZW_HD void dec_pred_whole(const u8* ws, int size, int mode, bool has_top, bool has_left, int bx, int by, i32* pr) {
  if (mode == 0) {  // predict_dcpred :183
    u32 sum = 0;
    int shf = size == 8 ? 2 : 3;
    if (has_left) { for (int y = 0; y < size; y++) sum += ws[(y + 1) * 32]; shf++; }
    if (has_top) { for (int x = 1; x <= size; x++) sum += ws[x]; shf++; }
    const i32 dc = (!has_left && !has_top) ? 128 : (i32)((sum + (1u << (shf - 1))) >> shf);
#pragma unroll
    for (int k = 0; k < 16; k++) pr[k] = dc;
    return;
  }
  const i32 p = ws[0];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const i32 t = ws[1 + bx * 4 + (k & 3)], l = ws[(1 + by * 4 + (k >> 2)) * 32];
    pr[k] = mode == 1 ? t : (mode == 2 ? l : clip255(l + t - p));
  }
}

// Reconstruction of the parsed macroblock into S.yws / S.uws / S.vws (intra_predict_luma :736, intra_predict_chroma :809,
// the inverse transforms of read_residual_data :1078-1166).  `tab` = ZW_PRED_TABLE_INIT.
template <class X>
ZW_HD void dec_reconstruct_mb(X& x, DecShared& S, int mbx, int mby, const u16 (*tab)[16]) {
  if (S.nflag[0] & 0x80) {
    x.run([&](int lane) {  // inverse WHT of Y2 -> the DCs of the 16 luma blocks (:1084-1088)
      if (lane != 0) return;
      i32 y2[16];
#pragma unroll
      for (int k = 0; k < 16; k++) y2[k] = S.y2[k];
      iwht4x4(y2);
#pragma unroll
      for (int k = 0; k < 16; k++) S.coef[k][0] = y2[k];
    });
  }
  const bool bpred = S.luma_mode == 4;
  x.run([&](int lane) {  // residuals of the 24 blocks; whole-block prediction + residue for chroma and non-B luma
    if (lane >= 24) return;
    i32 c[16];
#pragma unroll
    for (int k = 0; k < 16; k++) c[k] = S.coef[lane][k];
    const bool n = (S.nflag[lane] & 1) != 0;
    if (c[0] != 0 || n) {
      S.nzdct = 1;
      if (n) dec_idct4x4(c);
      else {
        const i32 dc = (c[0] + 4) >> 3;  // idct4x4_dc (transform.rs:13)
#pragma unroll
        for (int k = 0; k < 16; k++) c[k] = dc;
      }
    }
    if (lane < 16 && bpred) {
#pragma unroll
      for (int k = 0; k < 16; k++) S.coef[lane][k] = c[k];
      return;
    }
    u8* ws = lane < 16 ? S.yws : (lane < 20 ? S.uws : S.vws);
    const int bi = lane < 16 ? lane : (lane - 16) & 3;
    const int bx = lane < 16 ? bi & 3 : bi & 1, by = lane < 16 ? bi >> 2 : bi >> 1;
    i32 pr[16];
    dec_pred_whole(ws, lane < 16 ? 16 : 8, lane < 16 ? S.luma_mode : S.chroma_mode, mby != 0, mbx != 0, bx, by, pr);
#pragma unroll
    for (int k = 0; k < 16; k++) S.coef[lane][k] = clip255(pr[k] + c[k]);  // add_residue (prediction.rs:138); stored below
  });
  // (the whole-block predictors read row 0 / column 0 of the buffers only: writing the interior needs its own step so
  // that no lane overwrites what another lane's DC sum still reads -- column 0 and row 0 are never written)
  x.run([&](int lane) {
    if (lane >= 24 || (lane < 16 && bpred)) return;
    u8* ws = lane < 16 ? S.yws : (lane < 20 ? S.uws : S.vws);
    const int bi = lane < 16 ? lane : (lane - 16) & 3;
    const int bx = lane < 16 ? bi & 3 : bi & 1, by = lane < 16 ? bi >> 2 : bi >> 1;
#pragma unroll
    for (int k = 0; k < 16; k++) ws[(1 + by * 4 + (k >> 2)) * 32 + 1 + bx * 4 + (k & 3)] = (u8)S.coef[lane][k];
  });
  if (bpred) {
    // the reference walks the 16 sub-blocks in raster order; block (x, y) needs its left, top and top-right neighbours,
    // so blocks of equal x + 2y are independent: ten rounds
    for (int rd = 0; rd < 10; rd++) {
      x.run([&](int lane) {
        if (lane >= 16) return;
        const int sbx = lane & 3, sby = lane >> 2;
        if (sbx + 2 * sby != rd) return;
        const int x0 = 1 + 4 * sbx, y0 = 1 + 4 * sby;
        u8 e[13];
#pragma unroll
        for (int k = 0; k < 4; k++) e[k] = S.yws[(y0 + 3 - k) * 32 + x0 - 1];
#pragma unroll
        for (int k = 4; k < 13; k++) e[k] = S.yws[(y0 - 1) * 32 + x0 - 5 + k];
        const int mode = S.bpred[lane];
#pragma unroll
        for (int k = 0; k < 16; k++)
          S.yws[(y0 + (k >> 2)) * 32 + x0 + (k & 3)] = (u8)clip255(predict4_pixel(e, mode, k, tab) + S.coef[lane][k]);
      });
    }
  }
}

// calculate_filter_parameters (vp8.rs:1470-1524)
ZW_HD void dec_filter_params(const DecShared& S, u32 flags, int& fl, int& il, int& hev) {
  const int seg = (flags >> 5) & 3, luma_mode = flags & 7;
  i32 level = S.filter_level;
  if (level == 0) { fl = il = hev = 0; return; }
  if (S.segments_enabled) level = S.seg_delta ? level + S.seg_lf[seg] : (i32)S.seg_lf[seg];
  level = imin(imax(level, 0), 63);
  if (S.lf_adj) {
    level += S.ref_delta0;
    if (luma_mode == 4) level += S.mode_delta0;
  }
  fl = imin(imax(level, 0), 63);
  int interior = fl;
  if (S.sharpness > 0) {
    interior >>= S.sharpness > 4 ? 2 : 1;
    if (interior > 9 - S.sharpness) interior = 9 - S.sharpness;
  }
  if (interior == 0) interior = 1;
  il = interior;
  hev = fl >= 40 ? 2 : (fl >= 15 ? 1 : 0);
}

// ---- src/decoder/loop_filter.rs: one position of an edge; px = first pixel after the edge, s = distance of the taps ----
ZW_HD i32 lf_c(i32 v) { return imin(imax(v, -128), 127); }
ZW_HD i32 lf_u2s(u8 v) { return (i32)v - 128; }
ZW_HD u8 lf_s2u(i32 v) { return (u8)(lf_c(v) + 128); }
ZW_HD i32 lf_common_adjust(bool outer, u8* px, ptrdiff_t s) {  // :24
  const i32 p1 = lf_u2s(px[-2 * s]), p0 = lf_u2s(px[-s]), q0 = lf_u2s(px[0]), q1 = lf_u2s(px[s]);
  const i32 a0 = lf_c((outer ? lf_c(p1 - q1) : 0) + 3 * (q0 - p0));
  const i32 b = lf_c(a0 + 3) >> 3, a = lf_c(a0 + 4) >> 3;
  px[0] = lf_s2u(q0 - a);
  px[-s] = lf_s2u(p0 + b);
  return a;
}
ZW_HD bool lf_simple_threshold(i32 limit, const u8* px, ptrdiff_t s) {  // :70
  return iabs((i32)px[-s] - (i32)px[0]) * 2 + iabs((i32)px[-2 * s] - (i32)px[s]) / 2 <= limit;
}
ZW_HD bool lf_should_filter(i32 interior, i32 edge, const u8* px, ptrdiff_t s) {  // :90
  return lf_simple_threshold(edge, px, s) && iabs((i32)px[-4 * s] - (i32)px[-3 * s]) <= interior && iabs((i32)px[-3 * s] - (i32)px[-2 * s]) <= interior &&
         iabs((i32)px[-2 * s] - (i32)px[-s]) <= interior && iabs((i32)px[3 * s] - (i32)px[2 * s]) <= interior &&
         iabs((i32)px[2 * s] - (i32)px[s]) <= interior && iabs((i32)px[s] - (i32)px[0]) <= interior;
}
ZW_HD bool lf_hev(i32 thr, const u8* px, ptrdiff_t s) {  // :120
  return iabs((i32)px[-2 * s] - (i32)px[-s]) > thr || iabs((i32)px[s] - (i32)px[0]) > thr;
}
// kind 0: simple_segment (:132), 1: macroblock_filter (:190), 2: subblock_filter (:150)
ZW_HD void lf_apply(int kind, int hev_t, int interior, int edge, u8* px, ptrdiff_t s) {
  if (kind == 0) {
    if (lf_simple_threshold(edge, px, s)) lf_common_adjust(true, px, s);
    return;
  }
  if (!lf_should_filter(interior, edge, px, s)) return;
  const bool hv = lf_hev(hev_t, px, s);
  if (kind == 2) {
    const i32 a = (lf_common_adjust(hv, px, s) + 1) >> 1;
    if (!hv) {
      px[s] = lf_s2u(lf_u2s(px[s]) - a);
      px[-2 * s] = lf_s2u(lf_u2s(px[-2 * s]) + a);
    }
    return;
  }
  if (hv) { lf_common_adjust(true, px, s); return; }
  const i32 p2 = lf_u2s(px[-3 * s]), p1 = lf_u2s(px[-2 * s]), p0 = lf_u2s(px[-s]), q0 = lf_u2s(px[0]), q1 = lf_u2s(px[s]), q2 = lf_u2s(px[2 * s]);
  const i32 w = lf_c(lf_c(p1 - q1) + 3 * (q0 - p0));
  i32 a = lf_c((27 * w + 63) >> 7);
  px[0] = lf_s2u(q0 - a); px[-s] = lf_s2u(p0 + a);
  a = lf_c((18 * w + 63) >> 7);
  px[s] = lf_s2u(q1 - a); px[-2 * s] = lf_s2u(p1 + a);
  a = lf_c((9 * w + 63) >> 7);
  px[2 * s] = lf_s2u(q2 - a); px[-3 * s] = lf_s2u(p2 + a);
}

// filter_row_in_cache (vp8.rs:1172-1348) for one macroblock, in place on the frame: lanes 0..15 = the 16 luma rows /
// columns of an edge, 16..23 / 24..31 = the 8 chroma rows / columns of U / V (normal filter only).
template <class X>
ZW_HD void dec_filter_mb(X& x, const DecShared& S, u8* yp, u8* up, u8* vp, int mbx, int mby, u32 flags) {
  int fl, il, hev;
  dec_filter_params(S, flags, fl, il, hev);
  if (fl == 0) return;
  const int mbedge = (fl + 2) * 2 + il, sub = fl * 2 + il;
  const bool do_sub = (flags & 7) == 4 || (!((flags >> 7) & 1) && ((flags >> 8) & 1));
  const bool simple = S.filter_type != 0;
  const ptrdiff_t ys = (ptrdiff_t)S.mbw * 16, cs = (ptrdiff_t)S.mbw * 8;
  u8* Y = yp + (ptrdiff_t)mby * 16 * ys + mbx * 16;
  u8* U = up + (ptrdiff_t)mby * 8 * cs + mbx * 8;
  u8* V = vp + (ptrdiff_t)mby * 8 * cs + mbx * 8;
  // one step = one luma edge (+ the chroma edge that the reference filters right after it)
  auto vertical_edge = [&](int xoff, int kind, int limit, bool chroma) {  // lanes = rows
    x.run([&](int lane) {
      if (lane < 16) lf_apply(simple ? 0 : kind, hev, il, limit, Y + lane * ys + xoff, 1);
      else if (chroma && !simple) lf_apply(kind, hev, il, limit, (lane < 24 ? U : V) + (lane & 7) * cs + xoff / 2, 1);
    });
  };
  auto horizontal_edge = [&](int yoff, int kind, int limit, bool chroma) {  // lanes = columns
    x.run([&](int lane) {
      if (lane < 16) lf_apply(simple ? 0 : kind, hev, il, limit, Y + yoff * ys + lane, ys);
      else if (chroma && !simple) lf_apply(kind, hev, il, limit, (lane < 24 ? U : V) + (yoff / 2) * cs + (lane & 7), cs);
    });
  };
  if (mbx > 0) vertical_edge(0, 1, mbedge, true);
  if (do_sub) { vertical_edge(4, 2, sub, false); vertical_edge(8, 2, sub, true); vertical_edge(12, 2, sub, false); }
  if (mby > 0) horizontal_edge(0, 1, mbedge, true);
  if (do_sub) { horizontal_edge(4, 2, sub, false); horizontal_edge(8, 2, sub, true); horizontal_edge(12, 2, sub, false); }
}

// The whole frame of one image: header, macroblocks (parse on lane 0, reconstruction on all lanes), loop filter.
template <class X>
ZW_HD void dec_frame(X& x, DecShared& S, const DecParams& P, const DecImage& D, DecState& st, const u16 (*tab)[16]) {
  const u8* bytes = P.bytes;
  u64 off = D.data_off, len = D.data_len;
  x.run([&](int lane) {
    for (int i = lane; i < 1056; i += 32) (&S.probs[0][0][0][0])[i] = ZW_TAB(kCoeffProbs)[i];
    for (int i = lane; i < (int)D.mbw; i += 32) { P.topnz[D.col_off + i] = 0; P.topmodes[D.col_off + i] = 0; }
    if (lane == 0) S.status = ZWD_OK;
  });
  x.run([&](int lane) {
    if (lane != 0) return;
    DecState s0;
    s0 = st;
    s0.status = 0; s0.sse_rgb = 0;
    const int rc = dec_parse_header(S, bytes, off, len, D, s0);
    s0.status = (u32)rc;
    S.status = (u32)rc;
    st = s0;
  });
  if (S.status != ZWD_OK) return;
  const int mbw = (int)S.mbw, mbh = (int)S.mbh;
  const size_t ypitch = (size_t)mbw * 16, cpitch = (size_t)mbw * 8;
  u8* yp = P.planes + D.plane_off;
  u8* up = yp + ypitch * mbh * 16;
  u8* vp = up + cpitch * mbh * 8;
  for (int mby = 0; mby < mbh && S.status == ZWD_OK; mby++) {
    const int part = mby % (int)S.num_partitions;
    x.run([&](int lane) {
      if (lane == 0) { S.left_nz = 0; for (int i = 0; i < 4; i++) S.left_bpred[i] = 0; }
    });
    for (int mbx = 0; mbx < mbw; mbx++) {
      x.run([&](int lane) {  // clear the coefficient buffers; borders of the work buffers (create_border_luma / _chroma)
        for (int i = lane; i < 24 * 16; i += 32) (&S.coef[0][0])[i] = 0;
        if (lane < 16) S.y2[lane] = 0;
        if (lane < 24) S.nflag[lane] = 0;
        if (lane < 21) {  // luma row 0: corner, 16 above, 4 above-right
          u8 v;
          if (mby == 0) v = 127;
          else if (lane == 0) v = mbx == 0 ? (u8)129 : yp[(size_t)(mby * 16 - 1) * ypitch + mbx * 16 - 1];
          else if (lane <= 16 || mbx < mbw - 1) v = yp[(size_t)(mby * 16 - 1) * ypitch + mbx * 16 + lane - 1];
          else v = yp[(size_t)(mby * 16 - 1) * ypitch + mbx * 16 + 15];
          S.yws[lane] = v;
          if (lane >= 17) { S.yws[4 * 32 + lane] = v; S.yws[8 * 32 + lane] = v; S.yws[12 * 32 + lane] = v; }
        }
        if (lane < 16) S.yws[(1 + lane) * 32] = mbx == 0 ? (u8)129 : yp[(size_t)(mby * 16 + lane) * ypitch + mbx * 16 - 1];
        if (lane < 18) {  // chroma row 0 of U (lanes 0..8) and V (9..17)
          const int k = lane % 9;
          u8* ws = lane < 9 ? S.uws : S.vws;
          const u8* cp = lane < 9 ? up : vp;
          u8 v;
          if (mby == 0) v = 127;
          else if (k == 0) v = mbx == 0 ? (u8)129 : cp[(size_t)(mby * 8 - 1) * cpitch + mbx * 8 - 1];
          else v = cp[(size_t)(mby * 8 - 1) * cpitch + mbx * 8 + k - 1];
          ws[k] = v;
        }
        if (lane >= 16) {  // chroma column 0: U rows on lanes 16..23, V rows on 24..31
          const int r = lane & 7;
          u8* ws = lane < 24 ? S.uws : S.vws;
          const u8* cp = lane < 24 ? up : vp;
          ws[(1 + r) * 32] = mbx == 0 ? (u8)129 : cp[(size_t)(mby * 8 + r) * cpitch + mbx * 8 - 1];
        }
      });
      x.run([&](int lane) {
        if (lane != 0) return;
        const int rc = dec_parse_mb(S, bytes, P, D, mbx, part);
        if (rc != ZWD_OK) S.status = (u32)rc;
      });
      if (S.status != ZWD_OK) break;
      dec_reconstruct_mb(x, S, mbx, mby, tab);
      x.run([&](int lane) {  // store the macroblock; leave its modes for the filter sweep and the parity dump
        if (lane < 16) {
          u8* dst = yp + (size_t)(mby * 16 + lane) * ypitch + mbx * 16;
          const u8* srcp = &S.yws[(1 + lane) * 32 + 1];
#pragma unroll
          for (int k = 0; k < 16; k++) dst[k] = srcp[k];
        } else {
          const int r = lane & 7;
          u8* dst = (lane < 24 ? up : vp) + (size_t)(mby * 8 + r) * cpitch + mbx * 8;
          const u8* srcp = (lane < 24 ? S.uws : S.vws) + (1 + r) * 32 + 1;
#pragma unroll
          for (int k = 0; k < 8; k++) dst[k] = srcp[k];
        }
        if (lane == 0) {
          u32* mi = P.mbinfo + ((size_t)D.mb_off + (size_t)mby * mbw + mbx) * 4;
          u32 lo = 0, hi = 0;
          for (int i = 0; i < 8; i++) { lo |= (u32)S.bpred[i] << (4 * i); hi |= (u32)S.bpred[8 + i] << (4 * i); }
          mi[0] = (u32)S.luma_mode | ((u32)S.chroma_mode << 3) | ((u32)S.segment << 5) | ((u32)S.skipped << 7) | ((u32)S.nzdct << 8);
          mi[1] = lo; mi[2] = hi; mi[3] = 0;
        }
      });
    }
  }
  if (S.status != ZWD_OK) {
    x.run([&](int lane) { if (lane == 0) st.status = S.status; });
    return;
  }
  if (S.filter_level == 0) return;
  for (int mby = 0; mby < mbh; mby++)
    for (int mbx = 0; mbx < mbw; mbx++)
      dec_filter_mb(x, S, yp, up, vp, mbx, mby, P.mbinfo[((size_t)D.mb_off + (size_t)mby * mbw + mbx) * 4]);
}

// One output pixel: chroma upsampling + colour conversion.  fill_rgb_buffer_fancy (yuv.rs:82-157, rows :264-383,
// get_fancy_chroma_value :385), fill_rgb_buffer_simple (:402), conversion :36-78.
ZW_HD i32 dec_mulhi(u32 v, u32 coeff) { return (i32)((v * coeff) >> 8); }
ZW_HD u32 dec_clip6(i32 v) { return (u32)imin(imax(v >> 6, 0), 255); }
ZW_HD void dec_rgb_pixel(const u8* yp, const DecImage& D, int fancy, u32 row, u32 xx, u32* rgb) {
  const size_t ypitch = (size_t)D.mbw * 16, cpitch = (size_t)D.mbw * 8;
  const u8* up = yp + ypitch * D.mbh * 16;
  const u8* vp = up + cpitch * D.mbh * 8;
  const u32 cw = (D.width + 1) / 2, ch = (D.height + 1) / 2;
  const u32 c = xx / 2, c1 = row / 2;
  u32 u, v;
  if (fancy) {
    // the chroma sample nearest to the pixel weighs 9, its horizontal and vertical neighbours 3, the diagonal one 1;
    // at the image borders the neighbour is the sample itself
    const u32 nb = xx == 0 ? 0u : ((xx & 1) == 0 ? c - 1 : (c + 1 < cw ? c + 1 : c));
    const u32 c2 = row == 0 ? 0u : ((row & 1) ? (c1 + 1 < ch ? c1 + 1 : c1) : c1 - 1);
    const u8 *u1 = up + c1 * cpitch, *u2 = up + c2 * cpitch, *v1 = vp + c1 * cpitch, *v2 = vp + c2 * cpitch;
    u = (9u * u1[c] + 3u * u1[nb] + 3u * u2[c] + u2[nb] + 8u) >> 4;
    v = (9u * v1[c] + 3u * v1[nb] + 3u * v2[c] + v2[nb] + 8u) >> 4;
  } else {
    u = up[c1 * cpitch + c];
    v = vp[c1 * cpitch + c];
  }
  const u32 y = yp[row * ypitch + xx];
  rgb[0] = dec_clip6(dec_mulhi(y, 19077) + dec_mulhi(v, 26149) - 14234);
  rgb[1] = dec_clip6(dec_mulhi(y, 19077) - dec_mulhi(u, 6419) - dec_mulhi(v, 13320) + 8708);
  rgb[2] = dec_clip6(dec_mulhi(y, 19077) + dec_mulhi(u, 33050) - 17685);
}
// squared error of one decoded pixel against its source pixel (grey sources compare every channel with the grey value)
ZW_HD u32 dec_pixel_sse(const u8* s, u32 bpp, const u32* rgb) {
  const i32 sr = s[0], sg = bpp >= 3 ? s[1] : s[0], sb = bpp >= 3 ? s[2] : s[0];
  const i32 dr = (i32)rgb[0] - sr, dg = (i32)rgb[1] - sg, db = (i32)rgb[2] - sb;
  return (u32)(dr * dr + dg * dg + db * db);
}

#if defined(__CUDACC__)
__device__ const u16 d_dec_pred_tab[8][16] = ZW_PRED_TABLE_INIT;

struct DecWarpExec {
  int lane;
  template <class F>
  __device__ __forceinline__ void run(F&& f) {
    __syncwarp();
    f(lane);
    __syncwarp();
  }
};

constexpr int DEC_WARPS = 4;

// One warp per image: everything up to the filtered planes.
__global__ void __launch_bounds__(DEC_WARPS * 32) k_dec_frame(DecParams P) {
  __shared__ DecShared SH[DEC_WARPS];
  __shared__ u16 s_tab[8][16];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) (&s_tab[0][0])[i] = (&d_dec_pred_tab[0][0])[i];
  __syncthreads();
  const u32 img = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (img >= P.n_img) return;
  DecImage D = P.img[img];
  if (P.file_off) {  // verify after encode: the files sit in the encoder's output arena, placed by the device
    const ImageState es = P.enc_st[img];
    if (es.status != 0) { if ((threadIdx.x & 31) == 0) P.st[img].status = ZWD_CONTAINER; return; }
    D.data_off = P.file_off[img] + 20;
    D.data_len = es.vp8_bytes;
  }
  DecWarpExec X;
  X.lane = threadIdx.x & 31;
  dec_frame(X, SH[threadIdx.x >> 5], P, D, P.st[img], s_tab);
}

// Chroma upsampling + colour conversion (+ squared error against the source): one thread per pixel.
__global__ void __launch_bounds__(256) k_dec_rgb(DecParams P) {
  const u32 img = blockIdx.z;
  const DecImage D = P.img[img];
  const u32 row = blockIdx.y, xx = blockIdx.x * 256 + threadIdx.x;
  if (row >= D.height) return;  // uniform per block
  u32 sse = 0;
  if (xx < D.width && P.st[img].status == 0) {
    u32 rgb[3];
    dec_rgb_pixel(P.planes + D.plane_off, D, P.fancy, row, xx, rgb);
    if (P.rgb) {
      u8* o = P.rgb + D.rgb_off + ((size_t)row * D.width + xx) * 3;
      o[0] = (u8)rgb[0]; o[1] = (u8)rgb[1]; o[2] = (u8)rgb[2];
    }
    if (P.src && D.src_bpp) sse = dec_pixel_sse(P.src + D.src_off + ((size_t)row * D.width + xx) * D.src_bpp, D.src_bpp, rgb);
  }
  if (P.src) {
    __shared__ u32 s_sum[8];
    for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sse;
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 t = 0;
      for (int i = 0; i < 8; i++) t += s_sum[i];
      if (t) atomicAdd(reinterpret_cast<unsigned long long*>(&P.st[img].sse_rgb), (unsigned long long)t);
    }
  }
}
#endif  // __CUDACC__

}  // namespace zw
#endif
