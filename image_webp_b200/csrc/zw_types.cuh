// zw_types.cuh -- device-side data layout of one staged chunk of images (see DESIGN.md "HBM layout").
#ifndef ZW_TYPES_CUH
#define ZW_TYPES_CUH
#include <string.h>

#include "zw_cost.cuh"

namespace zw {

// Quantiser + lambda set of one quantiser index (Segment::init_matrices, src/common/types.rs:806-853;
// VP8Matrix::new, src/encoder/cost.rs:401-447).  Built on the host for all 128 indices.
struct SegParams {
  Matrix y1, y2, uv;
  u16 sharpen[16];  // y1 only
  u32 lambda_i4, lambda_i16, lambda_uv, lambda_mode;
  u32 lambda_trellis_i4, lambda_trellis_i16, tlambda;
  u32 uv_dc_zthresh;  // ((1<<17)-1-bias)/iq of the chroma DC entry (error diffusion, vp8.rs:600)
};

// Segment::init_matrices (src/common/types.rs:806-853) + VP8Matrix::new (src/encoder/cost.rs:401-447) for one quantiser
// index, on the host (zw_create builds all 128; tests/hostcheck builds the ones it needs).
inline u32 zw_umax(u32 a, u32 b) { return a > b ? a : b; }
inline u32 zw_umin(u32 a, u32 b) { return a < b ? a : b; }
inline Matrix make_matrix(u32 q_dc, u32 q_ac, int type) {
  static const u32 B[3][2] = {{96, 110}, {96, 108}, {110, 115}};
  Matrix m;
  m.q[0] = (u16)q_dc; m.q[1] = (u16)q_ac;
  for (int i = 0; i < 2; i++) {
    m.iq[i] = (u32)((1ull << 17) / (u64)m.q[i]);
    m.bias[i] = ((B[type][i] << 17) + 128) >> 8;
  }
  return m;
}
inline SegParams make_segparams(int idx) {
  SegParams s;
  memset(&s, 0, sizeof(s));
  const u32 ydc = (u32)host::kDcQuant[idx], yac = (u32)host::kAcQuant[idx];
  const u32 y2dc = ydc * 2, y2ac = zw_umax((u32)((int)yac * 155 / 100), 8);
  const u32 uvdc = ydc, uvac = yac;  // note: not clamped to 132 (Q18)
  s.y1 = make_matrix(ydc, yac, 0);
  s.y2 = make_matrix(y2dc, y2ac, 1);
  s.uv = make_matrix(uvdc, uvac, 2);
  for (int i = 0; i < 16; i++) s.sharpen[i] = (u16)(((u32)host::kFreqSharpening[i] * (u32)s.y1.q[i > 0]) >> 11);
  const u32 q_i4 = (ydc + 15 * yac + 8) >> 4, q_i16 = (y2dc + 15 * y2ac + 8) >> 4, q_uv = (uvdc + 15 * uvac + 8) >> 4;
  s.lambda_trellis_i4 = zw_umax((7 * q_i4 * q_i4) >> 3, 1);
  s.lambda_trellis_i16 = zw_umax((q_i16 * q_i16) >> 2, 1);
  s.lambda_i4 = zw_umax((3 * q_i4 * q_i4) >> 7, 1);
  s.lambda_i16 = zw_umax(3 * q_i16 * q_i16, 1);
  s.lambda_uv = zw_umax((3 * q_uv * q_uv) >> 6, 1);
  s.lambda_mode = zw_umax((q_i4 * q_i4) >> 7, 1);
  s.tlambda = (50u * q_i4) >> 5;
  s.uv_dc_zthresh = ((1u << 17) - 1 - s.uv.bias[0]) / s.uv.iq[0];
  return s;
}

// Per-macroblock record: the interface between the search passes, the statistics kernel and
// the tokeniser, and the P1MB / P2MB parity dump (832-byte POD, the layout tests/ compare against their CPU checker).
struct MbRecord {
  u8 ymode;   // 0 DC 1 V 2 H 3 TM 4 B_PRED
  u8 uvmode;  // 0 DC 1 V 2 H 3 TM
  u8 segment;
  u8 skip;
  u8 bmodes[16];
  u16 top_nz;   // incoming complexities: bit0 y2, 1..4 y[x], 5..6 u[x], 7..8 v[x]
  u16 left_nz;  // bit0 y2, 1..4 y[y], 5..6 u[y], 7..8 v[y]
  i8 derr_left[4];
  i8 derr_top[4];
  i16 levels[25][16];  // coded levels, zig-zag: [0] Y2, [1..16] Y, [17..20] U, [21..24] V
};
static_assert(sizeof(MbRecord) == 832, "MbRecord layout");

// Bottom row of a reconstructed macroblock, read by the row below (top / top-right borders).
struct MbBottom {
  u8 y[16];
  u8 u[8];
  u8 v[8];
};

// Host-filled description of one image of the chunk.
struct ImageDesc {
  u32 width, height, mbw, mbh;
  u32 bpp;       // 3 or 4
  u32 mb_off;    // index of the image's first macroblock in the chunk-wide MB arrays
  u32 row_off;   // index of its first macroblock row in the chunk-wide row arrays
  u32 use_segments;  // mbw*mbh >= 256 (vp8.rs:2481)
  u64 rgb_off;   // byte offset into the RGB arena (16-byte aligned)
  u64 y_off;     // byte offset of the padded Y plane in the plane arena; U follows, then V
};

// Device-computed placement of one image's variable-size data (k_layout, after the symbol count): the host never
// waits for the counts in the middle of a batch.
struct ImageLayout {
  u64 hdr_off;   // offset (in tokens) of the image's first-partition symbol stream
  u64 tok_off;   // offset (in tokens) of the image's token-partition symbol stream
  u64 part_off;  // byte offset of its coded partitions in the partition scratch: [p0 | p1]
  u64 out_off;   // byte offset of its finished file in the output arena (16-byte aligned; k_outscan)
  u32 p0_cap, p1_cap;  // byte capacities of the two coded partitions (7 bits per symbol bound)
};

// Chunk-wide totals of the same scan, checked against the capacities the host allocated up front.
struct ChunkTotals {
  u64 hdr_tokens, tok_tokens, part_bytes, out_bytes;
  u32 overflow;  // a stream arena is too small: the emit / code / assemble kernels do nothing, the host grows and re-runs them
  u32 segments;  // boolean-coder segments of all streams
};

// One boolean-coder segment (k_bc_*): what the passes hand to each other.
struct BcSegment {
  u32 cand[4];   // bitmap of the range states (range - 128 = bit index) the segment can possibly start in
  u64 start_bit; // total renormalisation shifts before the segment (absolute stream bit of its first addend)
  u64 tail;      // bottom << bit_num when the segment ends: still to be added to the bytes that follow
  u32 carries;   // carries that left the segment's own bytes towards earlier bytes
  u8 state;      // range - 1 the segment really starts with
  u8 pad[3];
};

// Device-written per-image state.
struct ImageState {
  u32 seg_count[4];
  u8 seg_qidx[4];
  i8 seg_delta[4];
  u8 tree_probs[3];
  u8 update_map;
  u8 seg_enabled;
  u8 skip_prob;
  u8 probs_updated;
  u8 centers[4];
  i32 mid_alpha;
  u32 n_skip1;       // skipped macroblocks in pass 1
  u32 hdr_tokens;    // tokens in the first-partition stream
  u32 tok_tokens;    // tokens in the token-partition stream
  u32 part0_bytes, part1_bytes;
  u32 vp8_bytes;    // frame tag .. end of the token partition (the VP8 payload)
  u32 file_bytes;   // RIFF header + payload + pad byte (api.rs:1325-1329), assembled on the device
  u32 status;   // 0 or a ZW_ERR_* code raised on the device
};
static_assert(sizeof(ImageState) <= 80, "ImageState is copied back per image; keep it small");

struct RowRef {
  u32 img;
  u32 mby;
};

// One boolean-coder input symbol: probability in the low byte, bit in bit 8.
typedef u16 Token;

// Everything the kernels need, passed by value.
struct ChunkParams {
  const ImageDesc* img;
  ImageLayout* lay;       // [n_img] device-computed stream / partition / output placement
  ChunkTotals* tot;
  u64 cap_hdr_tokens, cap_tok_tokens, cap_part_bytes;  // capacities of the symbol / partition arenas
  u64 cap_segments;       // capacity of the boolean-coder segment arrays
  u32* seg_off;           // [2 n_img + 1] first segment of every stream (token partitions, then first partitions)
  BcSegment* seg;         // [cap_segments]
  u32* seg_trans;         // [cap_segments][128] per candidate start state (by rank in cand): end state | total shift << 8
  ImageState* st;
  const RowRef* rows;     // ticket -> (image, mb row), ordered so that row y-1 precedes row y
  const SegParams* segtab;  // [128]
  const u8* segquant_lut;   // [128][255]: compute_segment_quant(base, alpha-127..127)
  u32 n_img, n_rows, n_mb;
  u32 method, base_qidx, do_trellis;
  u32 i4_modes;      // I4 candidates evaluated per sub-block: 0 (method <= 1: no I4), 3, 4 or 10 (vp8.rs:1836)
  u32 i4_always;     // method >= 5: I4 is tried for every macroblock (vp8.rs:2210-2231)
  u32 start_slack;   // a wavefront row starts once the row above is this many macroblocks ahead
  u8 filter_level;
  const u8* rgb;
  u8* planes;
  u32* alpha_hist;  // [n_img][256]
  u8* map256;       // [n_img][256] alpha -> segment (parity dump only)
  u8* alpha;     // [n_mb]
  u8* segmap;    // [n_mb]
  MbRecord* rec1;
  MbRecord* rec2;
  MbBottom* bottom;  // [n_mb]
  u16* nz_after;     // [n_mb] top complexity left behind by each MB
  u32* derr1;        // [n_mb] packed top_derr after pass 1
  u32* derr2;        // [n_mb] same, pass 2 (k_finish1 borrows it as scratch before)
  u32* c1info;       // [n_mb][2] pass-1 chroma chain -> k_finish1: left_derr after the MB, uv_mode | uvnz << 8
  u8* uvflags;       // [n_mb] pass-2 chroma has_coeffs bits (4 U | 4 V), k_chroma2 -> k_search<2>
  int* progress;     // [3][n_rows] macroblocks completed per row: pass-1 luma, pass-2 luma, pass-2 chroma
  u32* ticket;       // [4] work counters
  u32* rowstats;     // [n_rows][1056][2] per-row (total, ones)
  u32* stats;        // [n_img][1056] packed ProbaStats
  u8* probs;         // [n_img][1056] final token probabilities
  u16* lcost;        // [n_img][4*8*3*68]
  u32* mb_hdr_cnt;   // [n_mb+1] tokens per MB header (then exclusive scan)
  u32* mb_tok_cnt;   // [n_mb+1] tokens per MB residual
  u32* mb_lane_cnt;  // [n_mb][32] k_tokenize<0>: residual tokens of the lane's block | header-slot tokens << 16
  Token* hdr_tokens; // first-partition streams
  Token* tok_tokens; // token-partition streams
  u8* part_bytes;    // scratch for coded partitions
  u8* out;           // output arena
};

}  // namespace zw
#endif
