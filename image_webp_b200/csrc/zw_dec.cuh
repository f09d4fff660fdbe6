// zw_dec.cuh -- on-device VP8 key-frame DECODER, used as the batch verifier: every bitstream of a
// batch is parsed, reconstructed, loop-filtered, converted to RGB and scored against its source on the GPU.
//
// Reference (file:line under /root/reference):
//   boolean decoder        src/decoder/bit_reader.rs:29-250, :254-438         (rd_load / rd_bit)
//   frame header           src/decoder/vp8.rs:553-679 (+ :387, :421, :452, :506, :520)   (dec_parse_header)
//   macroblock header      src/decoder/vp8.rs:681-734                         (dec_parse_mb)
//   coefficients           src/decoder/vp8.rs:872-1058, residual data :1060-1170        (dec_read_coeffs, dec_parse_mb)
//   prediction + residue   src/decoder/vp8.rs:736-870; src/common/prediction.rs          (dec_recon_mb)
//   loop filter            src/decoder/vp8.rs:1172-1348, :1470-1524; src/decoder/loop_filter.rs      (dec_filter_mb)
//   YUV -> RGB             src/decoder/yuv.rs:36-78, :82-399 (bilinear), :402-530 (nearest)          (k_dec_rgb)
//
// Four kernels:
//   k_dec_parse   the two bitstream partitions of an image are strictly serial (every symbol's interval depends on the
//                 one before), so ONE WARP owns ONE IMAGE and lane 0 walks the bitstream; the warp only helps to move
//                 each macroblock's record (modes + 25 x 16 coded levels, zig-zag -- the layout the encoder's passes
//                 use, MbRecord) to HBM.  Latency bound by construction; the batch supplies the parallelism.
//   k_dec_rows<0> dequantisation, inverse transforms, prediction: a wavefront over macroblock ROWS (one warp per row,
//                 16 luma + 8 chroma 4x4 blocks on 24 lanes, rows of all images interleaved, row y waits for row y-1 to
//                 be one macroblock ahead -- the same ticket / progress-flag scheme as the encoder's k_search).
//   k_dec_rows<1> the loop filter, same wavefront (a macroblock's edges reach into its left, top and top-right
//                 neighbours), one lane per pixel row / column of an edge (16 luma + 8 + 8 chroma lanes); it runs after
//                 the whole frame is reconstructed because prediction reads UNFILTERED neighbours.
//   k_dec_rgb     chroma upsampling + colour conversion + squared error against the source, one thread per pixel.
//
// The per-macroblock work is written against an executor X (X::run(f) = "every lane calls f(lane), then the warp
// synchronises") and compiles for the host too: tests/hostcheck runs the same source lane by lane on the CPU and
// compares planes, modes and RGB with the decoder oracle before the code reaches a GPU.
#ifndef ZW_DEC_CUH
#define ZW_DEC_CUH
#include "zw_types.cuh"

#if !defined(__CUDACC__)  // host build (tests/hostcheck): the two CUDA vector types the record copies use
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = {x, y, z, w}; return v; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v = {x, y}; return v; }
#endif

namespace zw {

enum { ZWD_OK = 0, ZWD_BITSTREAM = 1, ZWD_UNSUPPORTED = 2, ZWD_MAGIC = 3, ZWD_COLORSPACE = 4, ZWD_TRUNCATED = 5, ZWD_CONTAINER = 6,
       ZWD_DIMENSIONS = 7 };

// Host-filled description of one file of a decode batch.
struct DecImage {
  u64 data_off;     // the VP8 frame (container already stripped) in the byte arena
  u32 data_len;
  u32 width, height, mbw, mbh;  // as the host read them from the frame header (or knows them from the source)
  u32 mb_off;       // first macroblock in the record / mbinfo arrays
  u32 row_off;      // first macroblock row in the progress arrays
  u32 src_bpp;      // verify: bytes per source pixel (1 L8, 2 La8, 3 Rgb8, 4 Rgba8), 0 = no source
  u64 plane_off;    // padded Y | U | V planes in the plane arena
  u64 rgb_off;      // decoded RGB (width * height * 3) in the RGB arena
  u64 src_off;      // verify: source pixels in the source arena
};

// Device-written per-image result + what the frame header leaves for the later kernels.
struct DecState {
  u32 status;
  u32 filter_type, filter_level, sharpness, num_partitions, segments_enabled, update_map, lf_adj, has_skip_prob, prob_skip_false,
      version, pixel_type;
  u64 sse_rgb;      // sum over the width x height x 3 colour samples of (decoded - source)^2
  i16 quant[4][6];  // per segment: ydc, yac, y2dc, y2ac, uvdc, uvac (vp8.rs:452-504)
  i8 seg_lf[4];     // per segment loop-filter level (absolute or delta)
  u8 seg_delta, pad[3];
  i32 ref_delta0, mode_delta0;
};

struct DecParams {
  const DecImage* img;
  DecState* st;
  u32 n_img, n_rows;
  int fancy;            // 1: bilinear chroma upsampling (the reference's default), 0: nearest
  const u8* bytes;      // file arena
  const u64* file_off;  // optional (verify after encode): device-side offsets of the files in `bytes` ...
  const ImageState* enc_st;  // ... and their sizes; data_off / data_len are then taken from here (+ 20-byte RIFF wrap)
  const RowRef* rows;   // ticket -> (image, macroblock row): row y of every image before row y + 1 of any
  MbRecord* rec;        // [n_mb] parsed macroblocks: modes, coded levels (zig-zag; [0] Y2, [1..16] Y, [17..24] U V),
                        //        top_nz | left_nz << 16 = the 24 "block has coefficients past its first position" bits
  u8* planes;
  u32* mbinfo;          // [n_mb][4]: flags (luma_mode | chroma_mode << 3 | segment << 5 | skipped << 7 | non_zero_dct << 8), bpred nibbles x2, 0
  int* progress;        // [2][n_rows] macroblocks completed per row: reconstruction, loop filter
  u32* ticket;          // [2]
  u8* rgb;              // may be null (verify only)
  const u8* src;        // may be null (decode only)
};

// ---- boolean decoder: libwebp's VP8GetBitAlt (bit_reader.rs:29-147), range - 1 stored.  The reference refills 56
// bits at a time on 64-bit hosts and 24 on 32-bit ones (BITS, :15-25); the decoded bits and the point at which `eof`
// is raised depend only on how many bits have been consumed, so the 24-bit form is used here: all state is 32-bit. ----
struct DecReader {
  u64 pos, end;
  u32 value, range;
  i32 bits;
  u32 eof;
};
ZW_HD int dec_clz(u32 v) {
#if defined(__CUDA_ARCH__)
  return __clz((int)v);
#else
  return __builtin_clz(v);
#endif
}
ZW_HD void rd_load(DecReader& r, const u8* bytes) {  // load_new_bytes :75 / load_final_bytes :59
  if (r.end - r.pos >= 3) {
    const u32 in = ((u32)bytes[r.pos] << 16) | ((u32)bytes[r.pos + 1] << 8) | (u32)bytes[r.pos + 2];
    r.value = in | (r.value << 24);
    r.bits += 24;
    r.pos += 3;
  } else if (r.pos < r.end) {
    r.bits += 8;
    r.value = (u32)bytes[r.pos] | (r.value << 8);
    r.pos++;
  } else if (!r.eof) {
    r.value <<= 8;
    r.bits += 8;
    r.eof = 1;
  } else r.bits = 0;
}
ZW_HD void rd_init(DecReader& r, const u8* bytes, u64 off, u64 len) {
  r.value = 0; r.range = 255 - 1; r.bits = -8; r.pos = off; r.end = off + len; r.eof = 0;
  rd_load(r, bytes);
}
ZW_HD int rd_bit(DecReader& r, const u8* bytes, u32 prob) {  // get_bit :120
  u32 range = r.range;
  if (r.bits < 0) rd_load(r, bytes);
  const int p = r.bits;
  const u32 split = (range * prob) >> 8;
  const u32 v = r.value >> p;
  const int bit = v > split;
  if (bit) { range -= split; r.value -= (split + 1) << p; }
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

constexpr int DEC_CTX_COLS = 256;  // macroblock columns whose top contexts live in shared memory (4096 px); wider: global

// What the parsing warp of an image keeps (one per warp in shared memory; a plain struct on the host).
struct DecParseShared {
  u8 ppos[4][17][3][11];  // token probabilities by coefficient POSITION (populate_probs_by_position, vp8.rs:405)
  u8 probs[4][8][3][11];  // ... by band, as the header updates them (types.rs:338, :387)
  DecReader rd[9];        // [0] first partition, [1 + p] token partition p
  alignas(16) MbRecord rec;  // the macroblock in flight
  u32 topmodes[DEC_CTX_COLS];  // bottom four sub-block modes of the row above
  u16 topnz[DEC_CTX_COLS];     // complexity it left behind (bit 0 y2, 1..4 y, 5..6 u, 7..8 v)
  u8 seg_probs[3];
  u8 segments_enabled, update_map, has_skip_prob, prob_skip_false, num_partitions;
  u8 left_bpred[4];
  u32 left_nz;
  u32 status;
};

// the trees of src/common/types.rs:191-205, :332 (a leaf is stored as -value)
#define ZW_DT_SEG {2, 4, 0, -1, -2, -3}
#define ZW_DT_YMODE {-4, 2, 4, 6, 0, -1, -2, -3}
#define ZW_DT_BMODE {0, 2, -1, 4, -2, 6, 8, 12, -3, 10, -5, -6, -4, 14, -7, 16, -8, -9}
#define ZW_DT_UV {0, 2, -1, 4, -2, -3}

// read_frame_header (vp8.rs:553-679), lane 0.  Leaves the readers and probabilities in S, quantisers and filter settings in st.
ZW_HD int dec_parse_header(DecParseShared& S, const u8* bytes, u64 off, u64 len, const DecImage& D, DecState& st) {
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
  u64 rpos = 10;
  if (len - rpos < first_size || first_size == 0) return ZWD_TRUNCATED;
  DecReader b;
  rd_init(b, bytes, off + rpos, first_size);
  rpos += first_size;
  const u32 color_space = rd_literal(b, bytes, 1);
  st.pixel_type = rd_literal(b, bytes, 1);
  if (color_space != 0) return ZWD_COLORSPACE;
  S.segments_enabled = (u8)rd_bit(b, bytes, 128);
  S.update_map = 0; st.seg_delta = 0;
  i8 seg_quant[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) st.seg_lf[i] = 0;
  for (int i = 0; i < 3; i++) S.seg_probs[i] = 255;
  if (S.segments_enabled) {  // read_segment_updates :520
    S.update_map = (u8)rd_bit(b, bytes, 128);
    if (rd_bit(b, bytes, 128)) {
      st.seg_delta = (u8)!rd_bit(b, bytes, 128);
      for (int i = 0; i < 4; i++) seg_quant[i] = (i8)rd_optional_signed(b, bytes, 7);
      for (int i = 0; i < 4; i++) st.seg_lf[i] = (i8)rd_optional_signed(b, bytes, 6);
    }
    if (S.update_map)
      for (int i = 0; i < 3; i++) S.seg_probs[i] = rd_bit(b, bytes, 128) ? (u8)rd_literal(b, bytes, 8) : (u8)255;
    if (b.eof) return ZWD_BITSTREAM;
  }
  st.filter_type = (u32)rd_bit(b, bytes, 128);
  st.filter_level = rd_literal(b, bytes, 6);
  st.sharpness = rd_literal(b, bytes, 3);
  st.lf_adj = (u32)rd_bit(b, bytes, 128);
  st.ref_delta0 = 0; st.mode_delta0 = 0;
  if (st.lf_adj) {  // read_loop_filter_adjustments :506 (only ref_delta[0] / mode_delta[0] matter for key frames)
    if (rd_bit(b, bytes, 128)) {
      for (int i = 0; i < 4; i++) { const i32 v = rd_optional_signed(b, bytes, 6); if (i == 0) st.ref_delta0 = v; }
      for (int i = 0; i < 4; i++) { const i32 v = rd_optional_signed(b, bytes, 6); if (i == 0) st.mode_delta0 = v; }
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
      const i32 base = S.segments_enabled ? (st.seg_delta ? (i32)seg_quant[i] + yac_abs : (i32)seg_quant[i]) : yac_abs;
      i32 v;
      st.quant[i][0] = ZW_TAB(kDcQuant)[imin(imax(base + ydc_d, 0), 127)];
      st.quant[i][1] = ZW_TAB(kAcQuant)[imin(imax(base, 0), 127)];
      st.quant[i][2] = (i16)(ZW_TAB(kDcQuant)[imin(imax(base + y2dc_d, 0), 127)] * 2);
      v = (i32)ZW_TAB(kAcQuant)[imin(imax(base + y2ac_d, 0), 127)] * 155 / 100;
      st.quant[i][3] = (i16)(v < 8 ? 8 : v);
      v = ZW_TAB(kDcQuant)[imin(imax(base + uvdc_d, 0), 127)];
      st.quant[i][4] = (i16)(v > 132 ? 132 : v);
      st.quant[i][5] = ZW_TAB(kAcQuant)[imin(imax(base + uvac_d, 0), 127)];
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
  st.num_partitions = S.num_partitions; st.segments_enabled = S.segments_enabled; st.update_map = S.update_map;
  st.has_skip_prob = S.has_skip_prob; st.prob_skip_false = S.prob_skip_false;
  return ZWD_OK;
}

// read_coefficients (vp8.rs:872-1058) without the dequantisation and the zig-zag: level of position n -> zz[n] (the
// reconstruction lanes undo both in parallel).  probs = the plane's [17][3][11] position table.  Returns 1 / 0
// (coefficients past `first` or not), -1 on a bitstream error.
ZW_HD int dec_read_coeffs(DecReader& r, const u8* bytes, const u8 (*probs)[3][11], int first, int ctx, i16* zz) {
  int n = first;
  const u8* prob = probs[n][ctx];
  while (n < 16) {
    if (!rd_bit(r, bytes, prob[0])) break;
    while (!rd_bit(r, bytes, prob[1])) {
      n++;
      if (n >= 16) return r.eof ? -1 : 1;
      prob = probs[n][0];
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
    zz[n] = (i16)v;
    n++;
    prob = probs[n][next_ctx];  // the table has 17 positions: no bounds test (populate_probs_by_position)
  }
  if (r.eof) return -1;
  return n > first;
}

// One macroblock: read_macroblock_header (:681) + read_residual_data (:1060) or the skip bookkeeping (:1546-1556).
// Lane 0.  S.rec is zero on entry.  tnz / topm: the column's top contexts (in and out).
ZW_HD int dec_parse_mb(DecParseShared& S, const u8* bytes, int part, u32& tnz_io, u32& topm_io) {
  const i8 T_SEG[6] = ZW_DT_SEG, T_YMODE[8] = ZW_DT_YMODE, T_BMODE[18] = ZW_DT_BMODE, T_UV[6] = ZW_DT_UV;
  MbRecord& R = S.rec;
  DecReader b = S.rd[0];
  u32 topm = topm_io;
  R.segment = (S.segments_enabled && S.update_map) ? (u8)rd_tree(b, bytes, T_SEG, S.seg_probs) : (u8)0;
  R.skip = S.has_skip_prob ? (u8)rd_bit(b, bytes, S.prob_skip_false) : (u8)0;
  R.ymode = (u8)rd_tree(b, bytes, T_YMODE, ZW_TAB(kKfYmodeProbs));
  if (R.ymode == 4) {
    for (int y = 0; y < 4; y++)
      for (int x = 0; x < 4; x++) {
        const int t = (topm >> (8 * x)) & 255, l = S.left_bpred[y];
        const u8 bm = (u8)rd_tree(b, bytes, T_BMODE, &ZW_TAB(kKfBmodeProbs)[(t * 10 + l) * 9]);
        R.bmodes[x + y * 4] = bm;
        topm = (topm & ~(255u << (8 * x))) | ((u32)bm << (8 * x));
        S.left_bpred[y] = bm;
      }
  } else {
    const u8 m = R.ymode == 0 ? 0 : (R.ymode == 1 ? 2 : (R.ymode == 2 ? 3 : 1));  // into_intra: DC, VE, HE, TM
    for (int i = 0; i < 4; i++) { R.bmodes[12 + i] = m; S.left_bpred[i] = m; }
    topm = m * 0x01010101u;
  }
  R.uvmode = (u8)rd_tree(b, bytes, T_UV, ZW_TAB(kKfUvModeProbs));
  topm_io = topm;
  S.rd[0] = b;
  if (b.eof) return ZWD_BITSTREAM;
  u32 tnz = tnz_io, lnz = S.left_nz, nflag = 0;
  if (R.skip) {
    if (R.ymode != 4) { tnz &= ~1u; lnz &= ~1u; }
    tnz &= 1u; lnz &= 1u;
  } else {
    DecReader r = S.rd[1 + part];
    int plane = R.ymode == 4 ? 3 : 1;  // Plane::YCoeff0 / Y2 (types.rs:48-57: YCoeff1 0, Y2 1, Chroma 2, YCoeff0 3)
    if (plane == 1) {
      const int ctx = (int)(tnz & 1) + (int)(lnz & 1);
      const int n = dec_read_coeffs(r, bytes, S.ppos[1], 0, ctx, R.levels[0]);
      if (n < 0) return ZWD_BITSTREAM;
      tnz = (tnz & ~1u) | (u32)n; lnz = (lnz & ~1u) | (u32)n;
      plane = 0;
    }
    const int first = plane == 0 ? 1 : 0;
    for (int y = 0; y < 4; y++) {
      u32 l = (lnz >> (1 + y)) & 1;
      for (int x = 0; x < 4; x++) {
        const int ctx = (int)((tnz >> (1 + x)) & 1) + (int)l;
        const int n = dec_read_coeffs(r, bytes, S.ppos[plane], first, ctx, R.levels[1 + x + y * 4]);
        if (n < 0) return ZWD_BITSTREAM;
        nflag |= (u32)n << (x + y * 4);
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
          const int n = dec_read_coeffs(r, bytes, S.ppos[2], 0, ctx, R.levels[1 + i]);
          if (n < 0) return ZWD_BITSTREAM;
          nflag |= (u32)n << i;
          l = (u32)n;
          tnz = (tnz & ~(1u << (j + x))) | ((u32)n << (j + x));
        }
        lnz = (lnz & ~(1u << (j + y))) | (l << (j + y));
      }
    }
    S.rd[1 + part] = r;
  }
  R.top_nz = (u16)(nflag & 0xffffu); R.left_nz = (u16)(nflag >> 16);
  tnz_io = tnz;
  S.left_nz = lnz;
  return ZWD_OK;
}

// The whole bitstream of one image -> P.rec (decode_frame_'s loop, vp8.rs:1531-1576, without the pixels).
// `topnz_g` / `topmodes_g`: global scratch for images wider than DEC_CTX_COLS macroblocks (else unused).
template <class X>
ZW_HD void dec_parse_frame(X& x, DecParseShared& S, const DecParams& P, const DecImage& D, u64 off, u64 len, DecState& st,
                           u16* topnz_g, u32* topmodes_g) {
  const u8* bytes = P.bytes;
  const bool wide = D.mbw > (u32)DEC_CTX_COLS;
  u16* topnz = wide ? topnz_g : S.topnz;
  u32* topmodes = wide ? topmodes_g : S.topmodes;
  x.run([&](int lane) {
    for (int i = lane; i < 1056; i += 32) (&S.probs[0][0][0][0])[i] = ZW_TAB(kCoeffProbs)[i];
    for (int i = lane; i < (int)D.mbw; i += 32) { topnz[i] = 0; topmodes[i] = 0; }
    for (int i = lane; i < (int)(sizeof(MbRecord) / 4); i += 32) reinterpret_cast<u32*>(&S.rec)[i] = 0;
    if (lane == 0) S.status = ZWD_OK;
  });
  x.run([&](int lane) {
    if (lane != 0) return;
    DecState s0 = st;
    s0.status = 0; s0.sse_rgb = 0;
    const int rc = dec_parse_header(S, bytes, off, len, D, s0);
    s0.status = (u32)rc;
    S.status = (u32)rc;
    st = s0;
  });
  if (S.status != ZWD_OK) return;
  x.run([&](int lane) {  // populate_probs_by_position (vp8.rs:405): position 16 is the look-ahead sentinel (band 7)
    for (int i = lane; i < 4 * 17 * 3; i += 32) {
      const int pl = i / 51, pos = (i / 3) % 17, ctx = i % 3;
      const int band = pos < 16 ? ZW_TAB(kCoeffBands)[pos] : 7;
      for (int t = 0; t < 11; t++) S.ppos[pl][pos][ctx][t] = S.probs[pl][band][ctx][t];
    }
  });
  const int mbw = (int)D.mbw, mbh = (int)D.mbh;
  for (int mby = 0; mby < mbh && S.status == ZWD_OK; mby++) {
    const int part = mby % (int)S.num_partitions;
    x.run([&](int lane) {
      if (lane == 0) { S.left_nz = 0; for (int i = 0; i < 4; i++) S.left_bpred[i] = 0; }
    });
    for (int mbx = 0; mbx < mbw; mbx++) {
      x.run([&](int lane) {
        if (lane != 0) return;
        u32 tnz = topnz[mbx], topm = topmodes[mbx];
        const int rc = dec_parse_mb(S, bytes, part, tnz, topm);
        topnz[mbx] = (u16)tnz; topmodes[mbx] = topm;
        if (rc != ZWD_OK) S.status = (u32)rc;
      });
      if (S.status != ZWD_OK) break;
      x.run([&](int lane) {  // the record goes to HBM in 16-byte pieces; the shared copy is cleared for the next one
        uint4* dst = reinterpret_cast<uint4*>(&P.rec[(size_t)D.mb_off + (size_t)mby * mbw + mbx]);
        uint4* srcp = reinterpret_cast<uint4*>(&S.rec);
        for (int i = lane; i < (int)(sizeof(MbRecord) / 16); i += 32) {
          dst[i] = srcp[i];
          srcp[i] = make_uint4(0, 0, 0, 0);
        }
      });
    }
  }
  if (S.status != ZWD_OK) x.run([&](int lane) { if (lane == 0) st.status = S.status; });
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

// What the reconstructing warp of a macroblock row keeps.
struct DecReconShared {
  i32 coef[24][16];  // residuals of the 24 blocks (natural order): 0..15 Y, 16..19 U, 20..23 V
  i32 y2[16];
  u8 yws[17 * 32];   // bordered work buffers (prediction.rs LUMA_STRIDE / CHROMA_STRIDE = 32)
  u8 uws[9 * 32];
  u8 vws[9 * 32];
  u8 bpred[16];
  u32 nzdct;
};

ZW_HD int dec_zigzag(int n) { return (int)((0xFEB7ADC963258410ull >> (4 * n)) & 15); }  // kZigzag as nibbles

// One macroblock from its record to pixels (read_residual_data's transforms :1078-1166, intra_predict_luma :736,
// intra_predict_chroma :809): planes and the mbinfo words are written; left / top borders are read from the planes
// (unfiltered: the filter kernel runs after the whole frame is reconstructed).  `tab` = ZW_PRED_TABLE_INIT.
template <class X>
ZW_HD void dec_recon_mb(X& x, DecReconShared& S, const DecParams& P, const DecImage& D, const DecState& st, int mbx, int mby,
                        const u16 (*tab)[16]) {
  const int mbw = (int)D.mbw, mbh = (int)D.mbh;
  const size_t ypitch = (size_t)mbw * 16, cpitch = (size_t)mbw * 8;
  u8* yp = P.planes + D.plane_off;
  u8* up = yp + ypitch * mbh * 16;
  u8* vp = up + cpitch * mbh * 8;
  const size_t mb = (size_t)D.mb_off + (size_t)mby * mbw + mbx;
  const MbRecord* R = &P.rec[mb];
  const u32 hdr = *reinterpret_cast<const u32*>(R);  // ymode | uvmode << 8 | segment << 16 | skip << 24
  const int luma_mode = hdr & 255, chroma_mode = (hdr >> 8) & 255, segment = (hdr >> 16) & 3;
  const bool skipped = (hdr >> 24) != 0, bpred = luma_mode == 4;
  const u32 nflag = (u32)R->top_nz | ((u32)R->left_nz << 16);
  const bool has_y2 = !bpred && !skipped;
  x.run([&](int lane) {
    // borders of the work buffers (create_border_luma / create_border_chroma, prediction.rs:15-126)
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
    if (lane == 0) S.nzdct = 0;
    if (lane < 16) S.bpred[lane] = R->bmodes[lane];
    // dequantisation + de-zig-zag of the block this lane owns; lane 24: Y2
    if (lane < 24 || (lane == 24 && has_y2)) {
      const i16* q = st.quant[segment];
      const int blk = lane == 24 ? 0 : 1 + lane;
      const i32 dcq = lane == 24 ? q[2] : (lane < 16 ? q[0] : q[4]), acq = lane == 24 ? q[3] : (lane < 16 ? q[1] : q[5]);
      i32* dst = lane == 24 ? S.y2 : S.coef[lane];
      if (skipped) {
#pragma unroll
        for (int n = 0; n < 16; n++) dst[n] = 0;
      } else {
        const uint4 a = reinterpret_cast<const uint4*>(R->levels[blk])[0], b4 = reinterpret_cast<const uint4*>(R->levels[blk])[1];
        const u32 w[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int n = 0; n < 16; n++) {
          const i32 lv = (i32)(i16)(w[n >> 1] >> (16 * (n & 1)));
          dst[dec_zigzag(n)] = lv * (n == 0 ? dcq : acq);
        }
      }
    }
  });
  if (has_y2) {
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
  x.run([&](int lane) {  // residuals of the 24 blocks; whole-block prediction + residue for chroma and non-B luma
    if (lane >= 24) return;
    i32 c[16];
#pragma unroll
    for (int k = 0; k < 16; k++) c[k] = S.coef[lane][k];
    const bool n = ((nflag >> lane) & 1) != 0;
    if (!skipped && (c[0] != 0 || n)) {
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
    dec_pred_whole(ws, lane < 16 ? 16 : 8, lane < 16 ? luma_mode : chroma_mode, mby != 0, mbx != 0, bx, by, pr);
#pragma unroll
    for (int k = 0; k < 16; k++) S.coef[lane][k] = clip255(pr[k] + c[k]);  // add_residue (prediction.rs:138); stored below
  });
  x.run([&](int lane) {  // (the whole-block predictors only read row 0 / column 0, which are never written)
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
  x.run([&](int lane) {  // store the macroblock; leave its modes for the filter kernel and the parity dump
    if (lane < 16) {
      u8* dst = yp + (size_t)(mby * 16 + lane) * ypitch + mbx * 16;
      const u8* srcp = &S.yws[(1 + lane) * 32 + 1];
      u32 w[4];
#pragma unroll
      for (int k = 0; k < 4; k++) w[k] = (u32)srcp[4 * k] | ((u32)srcp[4 * k + 1] << 8) | ((u32)srcp[4 * k + 2] << 16) | ((u32)srcp[4 * k + 3] << 24);
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      const int r = lane & 7;
      u8* dst = (lane < 24 ? up : vp) + (size_t)(mby * 8 + r) * cpitch + mbx * 8;
      const u8* srcp = (lane < 24 ? S.uws : S.vws) + (1 + r) * 32 + 1;
      u32 w[2];
#pragma unroll
      for (int k = 0; k < 2; k++) w[k] = (u32)srcp[4 * k] | ((u32)srcp[4 * k + 1] << 8) | ((u32)srcp[4 * k + 2] << 16) | ((u32)srcp[4 * k + 3] << 24);
      *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
    }
    if (lane == 0) {
      u32* mi = P.mbinfo + mb * 4;
      u32 lo = 0, hi = 0;
      for (int i = 0; i < 8; i++) { lo |= (u32)S.bpred[i] << (4 * i); hi |= (u32)S.bpred[8 + i] << (4 * i); }
      mi[0] = (u32)luma_mode | ((u32)chroma_mode << 3) | ((u32)segment << 5) | ((u32)skipped << 7) | (S.nzdct << 8);
      mi[1] = lo; mi[2] = hi; mi[3] = 0;
    }
  });
}

// calculate_filter_parameters (vp8.rs:1470-1524)
ZW_HD void dec_filter_params(const DecState& S, u32 flags, int& fl, int& il, int& hev) {
  const int seg = (flags >> 5) & 3, luma_mode = flags & 7;
  i32 level = (i32)S.filter_level;
  if (level == 0) { fl = il = hev = 0; return; }
  if (S.segments_enabled) level = S.seg_delta ? level + S.seg_lf[seg] : (i32)S.seg_lf[seg];
  level = imin(imax(level, 0), 63);
  if (S.lf_adj) {
    level += S.ref_delta0;
    if (luma_mode == 4) level += S.mode_delta0;
  }
  fl = imin(imax(level, 0), 63);
  int interior = fl;
  const int sharp = (int)S.sharpness;
  if (sharp > 0) {
    interior >>= sharp > 4 ? 2 : 1;
    if (interior > 9 - sharp) interior = 9 - sharp;
  }
  if (interior == 0) interior = 1;
  il = interior;
  hev = fl >= 40 ? 2 : (fl >= 15 ? 1 : 0);
}

// ---- src/decoder/loop_filter.rs: one position of an edge.  t[0..7] = p3 p2 p1 p0 q0 q1 q2 q3, filtered in registers;
//      returns a bit mask of the taps that changed ----
ZW_HD i32 lf_c(i32 v) { return imin(imax(v, -128), 127); }
ZW_HD i32 lf_u2s(i32 v) { return v - 128; }
ZW_HD i32 lf_s2u(i32 v) { return lf_c(v) + 128; }
ZW_HD i32 lf_common_adjust(bool outer, i32* t) {  // :24
  const i32 p1 = lf_u2s(t[2]), p0 = lf_u2s(t[3]), q0 = lf_u2s(t[4]), q1 = lf_u2s(t[5]);
  const i32 a0 = lf_c((outer ? lf_c(p1 - q1) : 0) + 3 * (q0 - p0));
  const i32 b = lf_c(a0 + 3) >> 3, a = lf_c(a0 + 4) >> 3;
  t[4] = lf_s2u(q0 - a);
  t[3] = lf_s2u(p0 + b);
  return a;
}
ZW_HD bool lf_simple_threshold(i32 limit, const i32* t) {  // :70
  return iabs(t[3] - t[4]) * 2 + iabs(t[2] - t[5]) / 2 <= limit;
}
ZW_HD bool lf_should_filter(i32 interior, i32 edge, const i32* t) {  // :90
  return lf_simple_threshold(edge, t) && iabs(t[0] - t[1]) <= interior && iabs(t[1] - t[2]) <= interior && iabs(t[2] - t[3]) <= interior &&
         iabs(t[7] - t[6]) <= interior && iabs(t[6] - t[5]) <= interior && iabs(t[5] - t[4]) <= interior;
}
ZW_HD bool lf_hev(i32 thr, const i32* t) { return iabs(t[2] - t[3]) > thr || iabs(t[5] - t[4]) > thr; }  // :120
// kind 0: simple_segment (:132), 1: macroblock_filter (:190), 2: subblock_filter (:150)
ZW_HD u32 lf_apply(int kind, int hev_t, int interior, int edge, i32* t) {
  if (kind == 0) {
    if (!lf_simple_threshold(edge, t)) return 0;
    lf_common_adjust(true, t);
    return 0x18;
  }
  if (!lf_should_filter(interior, edge, t)) return 0;
  const bool hv = lf_hev(hev_t, t);
  if (kind == 2) {
    const i32 a = (lf_common_adjust(hv, t) + 1) >> 1;
    if (hv) return 0x18;
    t[5] = lf_s2u(lf_u2s(t[5]) - a);
    t[2] = lf_s2u(lf_u2s(t[2]) + a);
    return 0x3c;
  }
  if (hv) { lf_common_adjust(true, t); return 0x18; }
  const i32 p2 = lf_u2s(t[1]), p1 = lf_u2s(t[2]), p0 = lf_u2s(t[3]), q0 = lf_u2s(t[4]), q1 = lf_u2s(t[5]), q2 = lf_u2s(t[6]);
  const i32 w = lf_c(lf_c(p1 - q1) + 3 * (q0 - p0));
  i32 a = lf_c((27 * w + 63) >> 7);
  t[4] = lf_s2u(q0 - a); t[3] = lf_s2u(p0 + a);
  a = lf_c((18 * w + 63) >> 7);
  t[5] = lf_s2u(q1 - a); t[2] = lf_s2u(p1 + a);
  a = lf_c((9 * w + 63) >> 7);
  t[6] = lf_s2u(q2 - a); t[1] = lf_s2u(p2 + a);
  return 0x7e;
}
// one position: load the eight taps around px (distance s), filter, store what changed
ZW_HD void lf_position(int kind, int hev_t, int interior, int edge, u8* px, ptrdiff_t s) {
  i32 t[8];
#pragma unroll
  for (int k = 0; k < 8; k++) t[k] = px[(k - 4) * s];
  const u32 m = lf_apply(kind, hev_t, interior, edge, t);
#pragma unroll
  for (int k = 1; k < 7; k++)
    if ((m >> k) & 1) px[(k - 4) * s] = (u8)t[k];
}

// filter_row_in_cache (vp8.rs:1172-1348) for one macroblock, in place on the frame: lanes 0..15 = the 16 luma rows /
// columns of an edge, 16..23 / 24..31 = the 8 chroma rows / columns of U / V (normal filter only).
template <class X>
ZW_HD void dec_filter_mb(X& x, const DecState& S, const DecImage& D, u8* planes, int mbx, int mby, u32 flags) {
  int fl, il, hev;
  dec_filter_params(S, flags, fl, il, hev);
  if (fl == 0) return;
  const int mbedge = (fl + 2) * 2 + il, sub = fl * 2 + il;
  const bool do_sub = (flags & 7) == 4 || (!((flags >> 7) & 1) && ((flags >> 8) & 1));
  const bool simple = S.filter_type != 0;
  const ptrdiff_t ys = (ptrdiff_t)D.mbw * 16, cs = (ptrdiff_t)D.mbw * 8;
  u8* yp = planes + D.plane_off;
  u8* up = yp + ys * D.mbh * 16;
  u8* vp = up + cs * D.mbh * 8;
  u8* Y = yp + (ptrdiff_t)mby * 16 * ys + mbx * 16;
  u8* U = up + (ptrdiff_t)mby * 8 * cs + mbx * 8;
  u8* V = vp + (ptrdiff_t)mby * 8 * cs + mbx * 8;
  // one step = one luma edge (+ the chroma edge of the same kind: the planes are independent)
  auto vertical_edge = [&](int xoff, int kind, int limit, bool chroma) {  // lanes = rows
    x.run([&](int lane) {
      if (lane < 16) lf_position(simple ? 0 : kind, hev, il, limit, Y + lane * ys + xoff, 1);
      else if (chroma && !simple) lf_position(kind, hev, il, limit, (lane < 24 ? U : V) + (lane & 7) * cs + xoff / 2, 1);
    });
  };
  auto horizontal_edge = [&](int yoff, int kind, int limit, bool chroma) {  // lanes = columns
    x.run([&](int lane) {
      if (lane < 16) lf_position(simple ? 0 : kind, hev, il, limit, Y + yoff * ys + lane, ys);
      else if (chroma && !simple) lf_position(kind, hev, il, limit, (lane < 24 ? U : V) + (yoff / 2) * cs + (lane & 7), cs);
    });
  };
  if (mbx > 0) vertical_edge(0, 1, mbedge, true);
  if (do_sub) { vertical_edge(4, 2, sub, false); vertical_edge(8, 2, sub, true); vertical_edge(12, 2, sub, false); }
  if (mby > 0) horizontal_edge(0, 1, mbedge, true);
  if (do_sub) { horizontal_edge(4, 2, sub, false); horizontal_edge(8, 2, sub, true); horizontal_edge(12, 2, sub, false); }
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

}  // namespace zw

#if defined(__CUDACC__)
#include "zw_search.cuh"  // progress-flag helpers (ld_flag / fence_acquire / publish_flag)

namespace zw {

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

constexpr int DEC_PARSE_WARPS = 2;  // warps (= images) per CTA of the parse kernel: spread the serial walks over all SMs
constexpr int DEC_ROW_WARPS = 8;    // warps (= macroblock rows in flight) per CTA of the wavefront kernels

// One warp per image: the bitstream walk.  topnz_g / topmodes_g: [n_img][1024] scratch for very wide images, or null.
__global__ void __launch_bounds__(DEC_PARSE_WARPS * 32) k_dec_parse(DecParams P, u16* topnz_g, u32* topmodes_g) {
  __shared__ DecParseShared SH[DEC_PARSE_WARPS];
  const u32 img = blockIdx.x * DEC_PARSE_WARPS + (threadIdx.x >> 5);
  if (img >= P.n_img) return;
  const DecImage D = P.img[img];
  u64 off = D.data_off, len = D.data_len;
  if (P.file_off) {  // verify after encode: the files sit in the encoder's output arena, placed by the device
    const ImageState es = P.enc_st[img];
    if (es.status != 0) { if ((threadIdx.x & 31) == 0) P.st[img].status = ZWD_CONTAINER; return; }
    off = P.file_off[img] + 20;
    len = es.vp8_bytes;
  }
  if (D.mbw > (u32)DEC_CTX_COLS && topnz_g == nullptr) { if ((threadIdx.x & 31) == 0) P.st[img].status = ZWD_DIMENSIONS; return; }
  DecWarpExec X;
  X.lane = threadIdx.x & 31;
  dec_parse_frame(X, SH[threadIdx.x >> 5], P, D, off, len, P.st[img], topnz_g ? topnz_g + (size_t)img * 1024 : nullptr,
                  topmodes_g ? topmodes_g + (size_t)img * 1024 : nullptr);
}

// Wavefront over macroblock rows: PHASE 0 reconstruction, PHASE 1 loop filter.  A warp takes rows by ticket (row y - 1 of
// an image always has a lower ticket than row y, so a waiting warp waits for a running one); macroblock x of row y
// needs x + 1 of row y - 1 (top-right neighbour; the last macroblock needs only its top neighbour).
template <int PHASE>
__global__ void __launch_bounds__(DEC_ROW_WARPS * 32) k_dec_rows(DecParams P) {
  __shared__ DecReconShared SH[PHASE == 0 ? DEC_ROW_WARPS : 1];
  __shared__ u16 s_tab[8][16];
  if (PHASE == 0) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) (&s_tab[0][0])[i] = (&d_dec_pred_tab[0][0])[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  DecWarpExec X;
  X.lane = lane;
  int* progress = P.progress + PHASE * P.n_rows;
  for (;;) {
    u32 t = 0;
    if (lane == 0) t = atomicAdd(&P.ticket[PHASE], 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= P.n_rows) return;
    const RowRef rr = P.rows[t];
    const DecImage D = P.img[rr.img];
    const DecState& st = P.st[rr.img];
    const int mby = (int)rr.mby, mbw = (int)D.mbw;
    const bool live = st.status == 0 && (PHASE == 0 || st.filter_level != 0);
    int seen = 0;
    for (int mbx = 0; mbx < mbw && live; mbx++) {
      if (mby > 0) {
        const int need = min(mbx + 2, mbw);
        if (seen < need) {
          if (lane == 0) {
            while ((seen = ld_flag(&progress[D.row_off + mby - 1])) < need) __nanosleep(64);
            fence_acquire();
          }
          seen = __shfl_sync(0xffffffffu, seen, 0);
        }
      }
      if constexpr (PHASE == 0) dec_recon_mb(X, SH[threadIdx.x >> 5], P, D, st, mbx, mby, s_tab);
      else dec_filter_mb(X, st, D, P.planes, mbx, mby, P.mbinfo[((size_t)D.mb_off + (size_t)mby * mbw + mbx) * 4]);
      __syncwarp();
      if (lane == 0) publish_flag(&progress[D.row_off + mby], mbx + 1);
    }
    if (!live && lane == 0) publish_flag(&progress[D.row_off + mby], mbw);  // nothing to do: do not hold the row below
  }
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

}  // namespace zw
#endif  // __CUDACC__
#endif
