// zw_capi.cu -- host side of the C ABI (include/zenwebp_b200.h): validation, HBM layout of a
// chunk, kernel sequencing, the pipeline of batches in flight, D2H.  Only the tiny f64
// quality->quantiser tables are computed on the host (the RIFF wrap is written by k_assemble);
// there is no CPU encode fallback.
//
// Reference (file:line under /root/reference):
//   quality_to_quant_index      src/encoder/vp8.rs:37-55       (+ fast_math.rs:15-43 cbrt)
//   compute_segment_quant       src/encoder/analysis.rs:1145-1174 (+ fast_math.rs:48-122 pow)
//   Segment::init_matrices      src/common/types.rs:806-853
//   VP8Matrix::new              src/encoder/cost.rs:401-447
//   compute_filter_level        src/encoder/cost.rs:271-294
//   RIFF container              src/encoder/api.rs:1224-1241, :1320-1329
// Compile the host part with -ffp-contract=off: Rust never fuses mul-add.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zenwebp_b200.h"
#include "zw_back.cuh"
#include "zw_dec.cuh"
#include "zw_front.cuh"
#include "zw_lossless.cuh"

using namespace zw;

namespace {

thread_local int g_last_error = 0;

// ---- f64 helpers (bit-exact restatement of the reference's no_std math) ----------------------
static inline u64 f64_bits(double x) { u64 b; memcpy(&b, &x, 8); return b; }
static inline double f64_from(u64 b) { double x; memcpy(&x, &b, 8); return x; }
static double h_cbrt(double x) {
  if (x == 0.0) return 0.0;
  double y = f64_from((f64_bits(x) / 3) + (1023ull * 2 / 3) * (1ull << 52));
  for (int i = 0; i < 4; i++) { double y2 = y * y; y = (2.0 * y + x / y2) / 3.0; }
  return y;
}
static double h_log2(double x) {
  const u64 bits = f64_bits(x);
  const i64 e = (i64)((bits >> 52) & 0x7FF) - 1023;
  const double m = f64_from((bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
  const double y = (m - 1.0) / (m + 1.0), y2 = y * y;
  const double poly = 2.8853900817779268 + y2 * (0.9617966939259756 + y2 * (0.5770780163555854 + y2 * (0.4121985831111324 + y2 * 0.3205988987531030)));
  return (double)e + y * poly;
}
static double h_exp2(double x) {
  x = x < -1022.0 ? -1022.0 : (x > 1023.0 ? 1023.0 : x);
  const i64 xi = x >= 0.0 ? (i64)x : (i64)x - 1;
  const double xf = x - (double)xi;
  const double LN2 = 0.693147180559945309417232121458176568;
  const double C1 = LN2, C2 = LN2 * LN2 / 2.0, C3 = LN2 * LN2 * LN2 / 6.0, C4 = LN2 * LN2 * LN2 * LN2 / 24.0,
               C5 = LN2 * LN2 * LN2 * LN2 * LN2 / 120.0;
  const double poly = 1.0 + xf * (C1 + xf * (C2 + xf * (C3 + xf * (C4 + xf * C5))));
  return poly * f64_from(((u64)(xi + 1023)) << 52);
}
static double h_pow(double x, double n) {
  if (x <= 0.0) return 0.0;
  if (x == 1.0 || n == 0.0) return 1.0;
  if (n == 1.0) return x;
  return h_exp2(n * h_log2(x));
}
static int quality_to_quant_index(int quality) {
  const double c = (double)quality / 100.0;
  const double lin = c < 0.75 ? c * (2.0 / 3.0) : 2.0 * c - 1.0;
  const int q = (int)(i64)(127.0 * (1.0 - h_cbrt(lin)) + 0.5);
  return std::min(std::max(q, 0), 127);
}
static int compute_segment_quant(int base_quant, int segment_alpha, int sns_strength) {
  const double amp = 0.9 * (double)sns_strength / 100.0 / 128.0;
  const double expn = 1.0 - amp * (double)segment_alpha;
  if (expn <= 0.0) return base_quant;
  const double c_base = 1.0 - ((double)base_quant / 127.0);
  const double c = h_pow(c_base, expn);
  const int q = (int)(127.0 * (1.0 - c));
  return std::min(std::max(q, 0), 127);
}
static int compute_filter_level(int quant_index) {  // sharpness 0, filter_strength 50 (vp8.rs:2417-2420)
  const u32 level0 = 5 * 50;
  const u32 qstep = (u32)(host::kAcTable[quant_index] >> 2) & 255;
  const u32 base = host::kLevelsFromDelta[0 * 64 + std::min<u32>(qstep, 63)];
  const u32 f = base * level0 / 256;
  return f < 2 ? 0 : (f > 63 ? 63 : (int)f);
}
// ---- token trees (RFC 6386 / src/common/types.rs:191-205, :332, :700-703) --------------------
static const i8 T_SEG[6] = {2, 4, 0, -1, -2, -3};
static const i8 T_YMODE[8] = {-4, 2, 4, 6, 0, -1, -2, -3};
static const i8 T_BMODE[18] = {0, 2, -1, 4, -2, 6, 8, 12, -3, 10, -5, -6, -4, 14, -7, 16, -8, -9};
static const i8 T_UV[6] = {0, 2, -1, 4, -2, -3};
static const i8 T_DCT[22] = {-11, 2, 0, 4, -1, 6, 8, 12, -2, 10, -3, -4, 14, 16, -5, -6, 18, 20, -7, -8, -9, -10};
static void walk_tree(const i8* tree, int node, u32 code, int len, TreeCodes& out) {
  for (int bit = 0; bit < 2; bit++) {
    const int nx = tree[node + bit];
    const u32 c = (code << 1) | (u32)bit;
    // leaves are stored as -value; value 0 is stored as 0, which is never a valid child index
    if (nx <= 0) { out.len[-nx] = (u8)(len + 1); out.code[-nx] = (u16)c; }
    else walk_tree(tree, nx, c, len + 1, out);
  }
}
static TokenTables make_token_tables() {
  TokenTables t;
  memset(&t, 0, sizeof(t));
  memcpy(t.tree_dct, T_DCT, sizeof(T_DCT)); memcpy(t.tree_ymode, T_YMODE, sizeof(T_YMODE));
  memcpy(t.tree_bmode, T_BMODE, sizeof(T_BMODE)); memcpy(t.tree_uv, T_UV, sizeof(T_UV)); memcpy(t.tree_seg, T_SEG, sizeof(T_SEG));
  walk_tree(T_DCT, 0, 0, 0, t.dct); walk_tree(T_YMODE, 0, 0, 0, t.ymode); walk_tree(T_BMODE, 0, 0, 0, t.bmode);
  walk_tree(T_UV, 0, 0, 0, t.uv); walk_tree(T_SEG, 0, 0, 0, t.seg);
  return t;
}

// ---- device / pinned-host buffers that only grow ----------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

#define CK(expr)                                                         \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) { g_last_error = ZW_ERR_CUDA + (int)_e; return g_last_error; } \
  } while (0)

}  // namespace

// One "lane": a stream with its own chunk buffers, i.e. one batch in flight.  A context owns `depth` lanes;
// zw_submit hands a batch to a free lane, whose H2D copy then runs under the kernels of the lane submitted
// before it and whose D2H copy runs under the kernels of the lane submitted after it.  Kernels of different
// lanes never overlap (measured: concurrent persistent kernels only slow each other down): a lane's first
// kernel waits for the previous lane's last one.
enum { LANE_FREE = 0, LANE_STAGED = 1, LANE_IN_FLIGHT = 2, LANE_SIZED = 3, LANE_DONE = 4 };
enum { EV_H2D0 = 0, EV_H2D1, EV_YUV0, EV_YUV1, EV_AN1, EV_P1, EV_C1, EV_ST0, EV_ST1, EV_C2, EV_P2, EV_TOK, EV_BC, EV_END, EV_START,
       EV_D2H0, EV_D2H1, EV_SIZES, EV_P1S, EV_F1S, EV_COUNT };
struct Lane {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;  // small chunks: the pass-1 chroma chains run here, beside the luma wavefront
  cudaEvent_t ev[EV_COUNT];
  bool ev_ok = false;
  const SegParams* segtab = nullptr;  // shared constant tables (owned by the context)
  const u8* lut = nullptr;
  DevBuf d_img, d_lay, d_tot, d_st, d_rows, d_rgb, d_planes, d_alpha_hist, d_map256, d_alpha, d_segmap, d_rec1, d_rec2, d_bottom, d_nz,
      d_derr1, d_derr2, d_c1, d_uvflags, d_progress, d_ticket, d_rowstats, d_stats, d_probs, d_lcost, d_hcnt, d_tcnt, d_lcnt, d_htok, d_ttok,
      d_part, d_out, d_outoff, d_segoff, d_seg, d_segtrans;
  PinBuf h_st, h_outoff, h_tot, h_arena;  // pinned mirrors: ImageState[n], out offsets[n+1], ChunkTotals, the finished files
  std::vector<ImageDesc> img;
  std::vector<RowRef> rows;
  std::vector<int> img_status;  // host-side validation result per image handed to this lane (0 = staged)
  std::vector<int> slot_of;     // image index within the lane -> position among staged (valid) images, or -1
  std::vector<u64> v_off;       // per input image: offset of its file (or payload) in h_arena
  std::vector<u32> v_len;
  std::vector<int32_t> v_status;
  size_t n_in = 0;
  u32 n_valid = 0, n_rows = 0, n_mb = 0, max_pw = 0, max_ph = 0, max_mb = 0;
  u64 layout_key = 0;
  u64 cap_h = 0, cap_t = 0, cap_p = 0, cap_seg = 0;  // capacities the phase-B arenas were launched with
  double tok_per_mb = 3.0 * 256;        // symbol-arena estimate (token partition), raised when a batch overflows it
  int state = LANE_FREE;
  int container = 1;
  int quality = -1, method = -1, base_qidx = 0;
  int search_blocks1 = 0, search_blocks2 = 0, chroma2_blocks = 0, sm_count = 0;
  int quad_blocks1 = 0, quad_blocks2 = 0;  // persistent grids of the quad (four lanes per row) luma kernels
  int quad_mode = -1;                      // -1 auto (by row count), 0 never, 1 always (ZW_QUAD)
  u32 start_slack = 0;  // measured: rows wait 1.1-1.4 % of their time at 1024 images; extra start slack only idles warps
  u64 launches = 0;
  u32 reruns = 0;
  dim3 tok_grid;
  ChunkParams P;
  zw_timing last;
};

// Buffers of the decoder / verifier entry points (zw_dec_host.inc); allocated on first use.
struct DecCtx {
  DevBuf d_img, d_st, d_bytes, d_planes, d_mbinfo, d_topnz, d_topmodes, d_rgb, d_src, d_rec, d_rows, d_progress;
  PinBuf h_st;
  cudaEvent_t ev[5];
  bool ev_ok = false;
  u32 last_n = 0;
  std::vector<DecImage> last_img;
  void release() {
    DevBuf* all[] = {&d_img, &d_st, &d_bytes, &d_planes, &d_mbinfo, &d_topnz, &d_topmodes, &d_rgb, &d_src, &d_rec, &d_rows, &d_progress};
    for (DevBuf* b : all) b->release();
    h_st.release();
    if (ev_ok) for (auto& e : ev) cudaEventDestroy(e);
    ev_ok = false;
  }
};

// Buffers of the lossless (VP8L) entry points (zw_lossless_host.inc); allocated on first use.  Two slots: two chunks of
// a batch are in flight at a time.
struct LlJob;
struct LlSlot {
  DevBuf d_img, d_st, d_src, d_res, d_desc, d_tile_last, d_tile_carry, d_tile_bits, d_tile_bitoff, d_hist, d_codes, d_hdr, d_out, d_outoff;
  PinBuf h_st, h_arena, h_outoff;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[9];
  bool ev_ok = false;
  std::vector<LlImage> img;
  std::vector<size_t> slot;  // job index (inside the chunk) of image k
  LlParams P;
  const LlJob* jobs = nullptr;
  zw_output* outs = nullptr;
  int container = 0;
  u32 ni = 0, n_tiles = 0;
  u64 src_bytes = 0, total = 0;
  void release() {
    DevBuf* all[] = {&d_img, &d_st, &d_src, &d_res, &d_desc, &d_tile_last, &d_tile_carry, &d_tile_bits, &d_tile_bitoff, &d_hist, &d_codes, &d_hdr,
                     &d_out, &d_outoff};
    if (stream) cudaStreamSynchronize(stream);
    for (DevBuf* b : all) b->release();
    h_st.release(); h_arena.release(); h_outoff.release();
    if (ev_ok) for (auto& e : ev) cudaEventDestroy(e);
    ev_ok = false;
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }
};
struct LlCtx {
  LlSlot slot[2];
  int last = -1;  // slot of the last chunk (zw_lossless_dump_stage)
  void release() { for (LlSlot& s : slot) s.release(); }
};

struct zw_ctx {
  int device = 0;
  int sm_count = 0;
  size_t budget = 0;
  DevBuf d_segtab, d_lut;
  std::vector<Lane*> lanes;
  Lane* prev = nullptr;  // lane whose kernels were launched last (the next lane's kernels wait for them)
  Lane* prev_copy = nullptr;  // lane whose H2D copy was enqueued last
  size_t n_staged = 0;   // split API (lane 0)
  bool staged = false, encoded = false;
  int dump_lane = -1;    // lane zw_dump_stage reads: the last chunk handed to the device
  zw_timing last;
  DecCtx dec;
  LlCtx ll;
};

static void fill_params(Lane* c) {
  ChunkParams& P = c->P;
  memset(&P, 0, sizeof(P));
  P.img = c->d_img.as<ImageDesc>(); P.lay = c->d_lay.as<ImageLayout>(); P.tot = c->d_tot.as<ChunkTotals>();
  P.cap_hdr_tokens = c->cap_h; P.cap_tok_tokens = c->cap_t; P.cap_part_bytes = c->cap_p; P.cap_segments = c->cap_seg;
  P.seg_off = c->d_segoff.as<u32>(); P.seg = c->d_seg.as<BcSegment>(); P.seg_trans = c->d_segtrans.as<u32>();
  P.st = c->d_st.as<ImageState>(); P.rows = c->d_rows.as<RowRef>();
  P.segtab = c->segtab; P.segquant_lut = c->lut;
  P.n_img = c->n_valid; P.n_rows = c->n_rows; P.n_mb = c->n_mb;
  P.rgb = c->d_rgb.as<u8>(); P.planes = c->d_planes.as<u8>();
  P.alpha_hist = c->d_alpha_hist.as<u32>(); P.map256 = c->d_map256.as<u8>();
  P.alpha = c->d_alpha.as<u8>(); P.segmap = c->d_segmap.as<u8>();
  P.rec1 = c->d_rec1.as<MbRecord>(); P.rec2 = c->d_rec2.as<MbRecord>(); P.bottom = c->d_bottom.as<MbBottom>();
  P.nz_after = c->d_nz.as<u16>(); P.derr1 = c->d_derr1.as<u32>(); P.derr2 = c->d_derr2.as<u32>(); P.c1info = c->d_c1.as<u32>(); P.uvflags = c->d_uvflags.as<u8>();
  P.progress = c->d_progress.as<int>(); P.ticket = c->d_ticket.as<u32>(); P.rowstats = c->d_rowstats.as<u32>();
  P.stats = c->d_stats.as<u32>(); P.probs = c->d_probs.as<u8>(); P.lcost = c->d_lcost.as<u16>();
  P.mb_hdr_cnt = c->d_hcnt.as<u32>(); P.mb_tok_cnt = c->d_tcnt.as<u32>(); P.mb_lane_cnt = c->d_lcnt.as<u32>();
  P.hdr_tokens = c->d_htok.as<Token>(); P.tok_tokens = c->d_ttok.as<Token>();
  P.part_bytes = c->d_part.as<u8>(); P.out = c->d_out.as<u8>();
  P.method = (u32)c->method; P.base_qidx = (u32)c->base_qidx; P.do_trellis = c->method >= 4;
  if (getenv("ZW_EXP_NOTRELLIS")) P.do_trellis = 0;  // timing experiment only: output is no longer the reference's
  P.i4_modes = c->method <= 1 ? 0u : (c->method <= 3 ? 3u : (c->method == 4 ? 4u : 10u));
  P.i4_always = c->method >= 5;
  P.filter_level = (u8)compute_filter_level(c->base_qidx);
  P.start_slack = c->start_slack;
}

static inline u32 color_bpp(u32 color) { return color + 1; }  // L8 1, La8 2, Rgb8 3, Rgba8 4 (api.rs:83-92)

static int validate_image(const zw_image& im) {
  if (im.color > ZW_COLOR_RGBA8) return ZW_ERR_INVALID_PARAM;
  if (im.width == 0 || im.height == 0 || im.width > 16383 || im.height > 16383) return ZW_ERR_INVALID_DIMENSIONS;
  if (im.data == nullptr) return im.len == 0 ? ZW_ERR_INVALID_BUFFER_SIZE : ZW_ERR_INVALID_PARAM;
  const u64 bpp = color_bpp(im.color);
  if ((u64)im.width * im.height * bpp != (u64)im.len) return ZW_ERR_INVALID_BUFFER_SIZE;
  return ZW_OK;
}

// Bytes of device memory one image needs (used to split a batch into chunks).
static size_t image_footprint(u32 w, u32 h, u32 bpp) {
  const size_t mbw = (w + 15) / 16, mbh = (h + 15) / 16, nmb = mbw * mbh;
  size_t b = (size_t)w * h * bpp + 64;        // RGB
  b += nmb * 384;                              // planes
  b += nmb * (2 * sizeof(MbRecord) + sizeof(MbBottom) + 2 + 2 + 8 + 8 + 8 + 128);
  b += mbh * (2112 * 4 + 8 + 8);               // row statistics, progress, row table
  b += 1056 * 7 + 6528 * 2 + 2048 + 9700 * 2;  // per-image tables, frame-header symbols
  b += nmb * 256 * 10 + nmb * 256 * 5;         // token streams (estimate: 5 symbols/px) + partitions + files
  return b;
}

static void lane_destroy(Lane* c) {
  if (!c) return;
  DevBuf* all[] = {&c->d_img, &c->d_lay, &c->d_tot, &c->d_st, &c->d_rows, &c->d_rgb, &c->d_planes, &c->d_alpha_hist, &c->d_map256, &c->d_alpha,
                   &c->d_segmap, &c->d_rec1, &c->d_rec2, &c->d_bottom, &c->d_nz, &c->d_derr1, &c->d_derr2, &c->d_c1, &c->d_uvflags, &c->d_progress,
                   &c->d_ticket, &c->d_rowstats, &c->d_stats, &c->d_probs, &c->d_lcost, &c->d_hcnt, &c->d_tcnt, &c->d_lcnt, &c->d_htok,
                   &c->d_ttok, &c->d_part, &c->d_out, &c->d_outoff, &c->d_segoff, &c->d_seg, &c->d_segtrans};
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (DevBuf* b : all) b->release();
  c->h_st.release(); c->h_outoff.release(); c->h_tot.release(); c->h_arena.release();
  if (c->ev_ok) for (auto& ev : c->ev) cudaEventDestroy(ev);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

static Lane* lane_create(zw_ctx* ctx, int warps_hint) {
  Lane* c = new Lane();
  c->device = ctx->device;
  c->segtab = ctx->d_segtab.as<SegParams>();
  c->lut = ctx->d_lut.as<u8>();
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
  if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return nullptr; }
  for (int i = 0; i < EV_COUNT; i++) {
    if (cudaEventCreate(&c->ev[i]) != cudaSuccess) {
      for (int k = 0; k < i; k++) cudaEventDestroy(c->ev[k]);
      cudaStreamDestroy(c->stream);
      delete c;
      return nullptr;
    }
  }
  c->ev_ok = true;
  if (c->d_ticket.reserve(64) != cudaSuccess || c->d_tot.reserve(sizeof(ChunkTotals)) != cudaSuccess ||
      c->h_tot.reserve(sizeof(ChunkTotals)) != cudaSuccess) { lane_destroy(c); return nullptr; }
  int b1 = 0, b2 = 0, b4 = 0;
  bool ok = true;
  ok &= cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b4, k_chroma2, SEARCH_WARPS * 32, sizeof(SearchShared)) == cudaSuccess;
  ok &= cudaFuncSetAttribute(k_search<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)search_smem_bytes(search_warps(1))) == cudaSuccess;
  ok &= cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_search<1>, search_warps(1) * 32, search_smem_bytes(search_warps(1))) == cudaSuccess;
  ok &= cudaFuncSetAttribute(k_search<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)search_smem_total(2)) == cudaSuccess;
  ok &= cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b2, k_search<2>, search_warps(2) * 32, search_smem_total(2)) == cudaSuccess;
  int bq1 = 0, bq2 = 0;
  ok &= cudaFuncSetAttribute(k_searchq<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)searchq_smem_bytes(1)) == cudaSuccess;
  ok &= cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bq1, k_searchq<1>, sq_warps(1) * 32, searchq_smem_bytes(1)) == cudaSuccess;
  ok &= cudaFuncSetAttribute(k_searchq<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)searchq_smem_bytes(2)) == cudaSuccess;
  ok &= cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bq2, k_searchq<2>, sq_warps(2) * 32, searchq_smem_bytes(2)) == cudaSuccess;
  if (!ok) { lane_destroy(c); return nullptr; }
  if (warps_hint > 0) { bq1 = std::min(bq1, std::max(1, warps_hint / sq_warps(1))); bq2 = std::min(bq2, std::max(1, warps_hint / sq_warps(2))); }
  c->quad_blocks1 = std::max(1, bq1) * ctx->sm_count;
  c->quad_blocks2 = std::max(1, bq2) * ctx->sm_count;
  if (const char* env = getenv("ZW_QUAD")) c->quad_mode = atoi(env);
  if (warps_hint > 0) {
    const int cap = std::max(1, warps_hint / SEARCH_WARPS);
    b1 = std::min(b1, std::max(1, warps_hint / search_warps(1))); b2 = std::min(b2, std::max(1, warps_hint / search_warps(2))); b4 = std::min(b4, cap);
  }
  if (const char* env = getenv("ZW_START_SLACK")) c->start_slack = (u32)std::max(0, atoi(env));
  c->sm_count = ctx->sm_count;
  c->chroma2_blocks = std::max(1, b4) * ctx->sm_count;
  c->search_blocks1 = std::max(1, b1) * ctx->sm_count;
  c->search_blocks2 = std::max(1, b2) * ctx->sm_count;
  return c;
}

// Validate + lay out + start the H2D copies of one lane's images (asynchronous on its stream).  The caller's
// buffers must stay valid and unchanged until the copy has finished (EV_H2D1).
static int lane_stage(Lane* c, const zw_image* imgs, size_t n, Lane* copy_after = nullptr) {
  c->state = LANE_FREE;
  c->n_in = n;
  c->img.clear(); c->img_status.assign(n, 0); c->slot_of.assign(n, -1);
  u64 rgb_bytes = 0, plane_bytes = 0, key = 1469598103934665603ull;
  u32 n_mb = 0, n_rows = 0, max_pw = 0, max_ph = 0, max_mb = 0, max_mbh = 0;
  for (size_t i = 0; i < n; i++) {
    const int v = validate_image(imgs[i]);
    c->img_status[i] = v;
    if (v != ZW_OK) continue;
    ImageDesc d;
    memset(&d, 0, sizeof(d));
    d.width = imgs[i].width; d.height = imgs[i].height;
    d.mbw = (d.width + 15) / 16; d.mbh = (d.height + 15) / 16;
    d.bpp = color_bpp(imgs[i].color);
    d.mb_off = n_mb; d.row_off = n_rows;
    d.use_segments = (d.mbw * d.mbh >= 256) ? 1 : 0;
    d.rgb_off = rgb_bytes; d.y_off = plane_bytes;
    rgb_bytes += ((u64)imgs[i].len + 15) & ~15ull;
    plane_bytes += (u64)d.mbw * d.mbh * 384;
    n_mb += d.mbw * d.mbh; n_rows += d.mbh;
    max_pw = std::max(max_pw, d.mbw * 16); max_ph = std::max(max_ph, d.mbh * 16);
    max_mb = std::max(max_mb, d.mbw * d.mbh); max_mbh = std::max(max_mbh, d.mbh);
    c->slot_of[i] = (int)c->img.size();
    c->img.push_back(d);
    key = (key ^ (((u64)d.width << 32) | ((u64)d.height << 8) | d.bpp)) * 1099511628211ull;
  }
  c->n_valid = (u32)c->img.size(); c->n_mb = n_mb; c->n_rows = n_rows;
  c->max_pw = max_pw; c->max_ph = max_ph; c->max_mb = max_mb;
  c->last = zw_timing();
  c->reruns = 0;
  if (c->n_valid == 0) { c->state = LANE_STAGED; return ZW_OK; }
  const u32 ni = c->n_valid;
  CK(c->d_img.reserve(ni * sizeof(ImageDesc))); CK(c->d_lay.reserve(ni * sizeof(ImageLayout))); CK(c->d_st.reserve(ni * sizeof(ImageState)));
  CK(c->d_rows.reserve((size_t)n_rows * sizeof(RowRef))); CK(c->d_rgb.reserve(rgb_bytes + 64)); CK(c->d_planes.reserve(plane_bytes + 64));
  CK(c->d_alpha_hist.reserve((size_t)ni * 1024)); CK(c->d_map256.reserve((size_t)ni * 256));
  CK(c->d_alpha.reserve(n_mb)); CK(c->d_segmap.reserve(n_mb));
  CK(c->d_rec1.reserve((size_t)n_mb * sizeof(MbRecord))); CK(c->d_rec2.reserve((size_t)n_mb * sizeof(MbRecord)));
  CK(c->d_bottom.reserve((size_t)n_mb * sizeof(MbBottom))); CK(c->d_nz.reserve((size_t)n_mb * 2));
  CK(c->d_derr1.reserve((size_t)n_mb * 4)); CK(c->d_derr2.reserve((size_t)n_mb * 4)); CK(c->d_c1.reserve((size_t)n_mb * 8)); CK(c->d_uvflags.reserve(n_mb));
  CK(c->d_progress.reserve((size_t)n_rows * 3 * sizeof(int))); CK(c->d_rowstats.reserve((size_t)n_rows * 2112 * 4));
  CK(c->d_stats.reserve((size_t)ni * 1056 * 4)); CK(c->d_probs.reserve((size_t)ni * 1056)); CK(c->d_lcost.reserve((size_t)ni * 6528 * 2));
  CK(c->d_hcnt.reserve(((size_t)n_mb + 1) * 4)); CK(c->d_tcnt.reserve(((size_t)n_mb + 1) * 4)); CK(c->d_lcnt.reserve((size_t)n_mb * 128));
  CK(c->d_outoff.reserve(((size_t)ni + 1) * 8));
  CK(c->h_st.reserve(ni * sizeof(ImageState))); CK(c->h_outoff.reserve(((size_t)ni + 1) * 8));
  // ticket order: macroblock row y of every image before row y+1 of any image ("many images
  // interleaved"): each image has at most a couple of rows in flight, so rows rarely wait.
  if (key != c->layout_key || c->rows.size() != n_rows) {
    c->rows.resize(n_rows);
    size_t k = 0;
    for (u32 y = 0; y < max_mbh; y++)
      for (u32 i = 0; i < ni; i++)
        if (y < c->img[i].mbh) { c->rows[k].img = i; c->rows[k].mby = y; k++; }
    CK(cudaMemcpyAsync(c->d_rows.p, c->rows.data(), (size_t)n_rows * sizeof(RowRef), cudaMemcpyHostToDevice, c->stream));
    c->layout_key = key;
  }
  CK(cudaMemcpyAsync(c->d_img.p, c->img.data(), ni * sizeof(ImageDesc), cudaMemcpyHostToDevice, c->stream));
  // copies of different lanes go over the link one after the other: sharing it would only delay the batch submitted
  // first (its kernels could not start) without finishing the later one any sooner
  if (copy_after && copy_after != c && copy_after->n_valid) CK(cudaStreamWaitEvent(c->stream, copy_after->ev[EV_H2D1], 0));
  CK(cudaEventRecord(c->ev[EV_H2D0], c->stream));
  {  // one copy per run of images that are contiguous in host memory (and need no alignment gap on the device)
    const u8* run_src = nullptr;
    u64 run_dst = 0, run_len = 0;
    for (size_t i = 0; i < n; i++) {
      if (c->slot_of[i] < 0) continue;
      const ImageDesc& d = c->img[c->slot_of[i]];
      if (run_len && imgs[i].data == run_src + run_len && d.rgb_off == run_dst + run_len) { run_len += imgs[i].len; continue; }
      if (run_len) CK(cudaMemcpyAsync(c->d_rgb.as<u8>() + run_dst, run_src, run_len, cudaMemcpyHostToDevice, c->stream));
      run_src = imgs[i].data; run_dst = d.rgb_off; run_len = imgs[i].len;
    }
    if (run_len) CK(cudaMemcpyAsync(c->d_rgb.as<u8>() + run_dst, run_src, run_len, cudaMemcpyHostToDevice, c->stream));
  }
  CK(cudaEventRecord(c->ev[EV_H2D1], c->stream));
  c->last.h2d_bytes = rgb_bytes;
  for (const ImageDesc& d : c->img) c->last.pixels += (u64)d.width * d.height;
  c->state = LANE_STAGED;
  return ZW_OK;
}

// Phase B: place the streams (k_layout), emit + code + assemble, then start the D2H of the per-image results.
// Everything is asynchronous; the arenas were sized before the symbol counts were known (see lane_sync_sizes).
static int lane_launch_b(Lane* c) {
  const u32 ni = c->n_valid;
  cudaStream_t s = c->stream;
  // first-partition symbols have a hard bound (MB header <= 122 symbols, frame header <= 9620); the token partition
  // is sized by estimate (raised on overflow); a symbol codes to at most 7 bits
  c->cap_h = (u64)c->n_mb * 128 + (u64)ni * 9700;
  c->cap_t = std::max(c->cap_t, (u64)((double)c->n_mb * c->tok_per_mb) + (u64)ni * 8);
  c->cap_p = ((c->cap_h + c->cap_t) * 7) / 8 + (u64)ni * 48;
  CK(c->d_htok.reserve(c->cap_h * sizeof(Token) + 64)); CK(c->d_ttok.reserve(c->cap_t * sizeof(Token) + 64));
  CK(c->d_part.reserve(c->cap_p + 64)); CK(c->d_out.reserve(c->cap_p + (u64)ni * 64 + 64));
  c->cap_seg = (c->cap_h + c->cap_t) / BC_SEG + 2ull * ni + 8;
  CK(c->d_segoff.reserve((2ull * ni + 1) * 4)); CK(c->d_seg.reserve(c->cap_seg * sizeof(BcSegment))); CK(c->d_segtrans.reserve(c->cap_seg * 128 * 4));
  fill_params(c);
  ChunkParams& P = c->P;
  k_layout<<<1, 1024, 0, s>>>(P);
  k_tokenize<1><<<c->tok_grid, TOK_WARPS * 32, 0, s>>>(P);
  k_frame_header<<<(ni + 63) / 64, 64, 0, s>>>(P);
  c->launches += 3;
  CK(cudaEventRecord(c->ev[EV_TOK], s));
  {  // boolean coder: candidate start states -> transitions -> resolve -> code (one lane per segment) -> fix up
    const u64 wide = (u64)c->sm_count * 16;
    k_bc_cands<<<(unsigned)std::min<u64>(c->cap_seg, wide), 128, 0, s>>>(P);
    k_bc_trans<<<(unsigned)std::min<u64>((c->cap_seg * 4 + 127) / 128, wide), 128, 0, s>>>(P);
    k_bc_resolve<<<(2 * ni + 127) / 128, 128, 0, s>>>(P);
    k_bc_code<<<(unsigned)std::min<u64>((c->cap_seg + 63) / 64, wide * 2), 64, 0, s>>>(P);
    k_bc_fix<<<(2 * ni + 127) / 128, 128, 0, s>>>(P);
    c->launches += 5;
  }
  CK(cudaEventRecord(c->ev[EV_BC], s));
  k_outscan<<<1, 32, 0, s>>>(P, c->d_outoff.as<u64>());
  k_assemble<<<ni, 256, 0, s>>>(P, c->d_outoff.as<u64>());
  c->launches += 2;
  CK(cudaEventRecord(c->ev[EV_END], s));
  CK(cudaMemcpyAsync(c->h_st.p, c->d_st.p, ni * sizeof(ImageState), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(c->h_outoff.p, c->d_outoff.p, ((size_t)ni + 1) * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(c->h_tot.p, c->d_tot.p, sizeof(ChunkTotals), cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(c->ev[EV_SIZES], s));
  return ZW_OK;
}

// Every kernel of one batch, asynchronous on the lane's stream (22 launches).
static int lane_launch(Lane* c, int quality, int method, Lane* after) {
  c->launches = 0;
  c->quality = quality; c->method = method; c->base_qidx = quality_to_quant_index(quality);
  if (c->n_valid == 0) { c->state = LANE_IN_FLIGHT; return ZW_OK; }
  const u32 ni = c->n_valid;
  cudaStream_t s = c->stream;
  // Kernels of different batches never overlap: a batch's first kernel waits for the last kernel of the batch before it.
  // (Measured: letting the front end of batch N+1 -- colour conversion, analysis, segments, the pass-1 chroma chains --
  // run under the tokeniser / boolean coder of batch N moves 11 ms of device time per batch but gains 0.5 ms per step:
  // those kernels fill the machine, and the latency-bound chains starve under them as they do under the wavefront.)
  if (after && after != c && after->n_valid) CK(cudaStreamWaitEvent(s, after->ev[EV_END], 0));
  fill_params(c);
  ChunkParams& P = c->P;
  bool use_quads = false;
  CK(cudaEventRecord(c->ev[EV_START], s));
  CK(cudaMemsetAsync(c->d_st.p, 0, ni * sizeof(ImageState), s));
  CK(cudaMemsetAsync(c->d_alpha_hist.p, 0, (size_t)ni * 1024, s));
  CK(cudaMemsetAsync(c->d_progress.p, 0, (size_t)c->n_rows * 3 * sizeof(int), s));
  CK(cudaMemsetAsync(c->d_ticket.p, 0, 64, s));
  CK(cudaEventRecord(c->ev[EV_YUV0], s));  // yuv_ms times k_yuv alone; device_total_ms starts at EV_START
  {  // (1) RGB -> YUV420
    const int rows_per_cta = YUV_ROWPAIRS * YUV_STEPS * 2;
    dim3 grid((c->max_pw + YUV_TILE_W - 1) / YUV_TILE_W, (c->max_ph + rows_per_cta - 1) / rows_per_cta, ni);
    dim3 block(YUV_THREADS, YUV_ROWPAIRS);
    const size_t sm = (size_t)YUV_ROWPAIRS * 4 * YUV_ROW_SLOTS * 16;  // two buffers of two rows per warp
    k_yuv<<<grid, block, sm, s>>>(P);
    c->launches++;
  }
  CK(cudaEventRecord(c->ev[EV_YUV1], s));
  {  // (2) analysis + segments
    bool any_seg = false;
    for (const ImageDesc& d : c->img) any_seg |= d.use_segments != 0;
    if (any_seg) {
      dim3 grid((c->max_mb + AN_WARPS - 1) / AN_WARPS, ni);
      k_analysis<<<grid, AN_WARPS * 32, 0, s>>>(P);
      c->launches++;
    }
    k_segments<<<ni, 256, 0, s>>>(P);
    c->launches++;
  }
  CK(cudaEventRecord(c->ev[EV_AN1], s));
  {  // (3) pass 1: the per-image chroma chains (independent of luma: disjoint words of the records), then the luma
     // wavefront, then the bookkeeping.  (Running the chains on a side stream UNDER the wavefront was measured: they
     // starve -- 43 ms instead of 8.8 ms, instruction-cache contention with the wavefront's code -- so the kernels stay
     // back to back.)
    // Chunks too small to fill the GPU (single images): the chains run on a side stream BESIDE the luma wavefront, which
    // then occupies a fraction of the SMs and starves nobody (4096 x 4096: 21 ms of luma under the 155 ms chain).
    const bool quad_fill = c->n_rows >= (u32)(c->sm_count * 64 * 2);
    const bool beside = c->n_rows < (u32)(c->sm_count * 8) && !getenv("ZW_NO_SIDE");
    const int g3 = (int)(((u64)ni + SEARCH_WARPS - 1) / SEARCH_WARPS);
    if (beside) {
      CK(cudaStreamWaitEvent(c->side, c->ev[EV_AN1], 0));
      k_chroma1<<<g3, SEARCH_WARPS * 32, sizeof(SearchShared), c->side>>>(P);
      CK(cudaEventRecord(c->ev[EV_C1], c->side));
    } else {
      k_chroma1<<<g3, SEARCH_WARPS * 32, sizeof(SearchShared), s>>>(P);
      CK(cudaEventRecord(c->ev[EV_C1], s));
    }
    CK(cudaEventRecord(c->ev[EV_P1S], s));
    // Batches with an I4 search: four lanes per macroblock row (k_searchq), eight rows per warp -- the I4 candidates run
    // lane-private (measured on 1024 x 768x512: pass 1 20.7 -> 17.5 ms at method 4, 31.9 -> 24.1 ms at method 6).  One warp
    // per row (k_search) stays for: too few rows to fill the GPU with quads (single images); methods 0 / 1 (no I4: the
    // warp kernel's lane-private I16 layout is already dense, 7.2 vs 8.3 ms); pass 2 with trellis (a macroblock offers at
    // most two independent trellis blocks, so half a quad idles: 39 vs 54 ms).  ZW_QUAD=0 / 1 / 2 forces never / both / pass 1.
    use_quads = c->quad_mode < 0 ? (quad_fill && P.i4_modes > 0) : c->quad_mode != 0;
    if (use_quads) {
      const int g1 = (int)std::min<u64>((u64)c->quad_blocks1, ((u64)c->n_rows + sq_quads(1) - 1) / sq_quads(1));
      k_searchq<1><<<g1, sq_warps(1) * 32, searchq_smem_bytes(1), s>>>(P);
    } else {
      const int w1 = search_warps(1);
      const int g1 = (int)std::min<u64>((u64)c->search_blocks1, ((u64)c->n_rows + w1 - 1) / w1);
      k_search<1><<<g1, w1 * 32, search_smem_bytes(w1), s>>>(P);
    }
    CK(cudaEventRecord(c->ev[EV_P1], s));
    if (beside) CK(cudaStreamWaitEvent(s, c->ev[EV_C1], 0));
    CK(cudaEventRecord(c->ev[EV_F1S], s));
    k_finish1<<<ni, 256, 0, s>>>(P);
    c->launches += 3;
  }
  CK(cudaEventRecord(c->ev[EV_ST0], s));
  {  // (4) token statistics -> probabilities, level costs, skip probability
    k_rowstats<<<(c->n_rows + STAT_WARPS - 1) / STAT_WARPS, STAT_WARPS * 32, 0, s>>>(P);
    k_probs<<<ni, 256, 0, s>>>(P);
    c->launches += 2;
  }
  CK(cudaEventRecord(c->ev[EV_ST1], s));
  {  // (3') pass 2: chroma wavefront first (independent of luma), then the luma wavefront
    const int g4 = (int)std::min<u64>((u64)c->chroma2_blocks, ((u64)c->n_rows + SEARCH_WARPS - 1) / SEARCH_WARPS);
    k_chroma2<<<g4, SEARCH_WARPS * 32, sizeof(SearchShared), s>>>(P);
    CK(cudaEventRecord(c->ev[EV_C2], s));
    if (c->quad_mode < 0 ? (use_quads && !P.do_trellis) : c->quad_mode == 1) {
      const int g2 = (int)std::min<u64>((u64)c->quad_blocks2, ((u64)c->n_rows + sq_quads(2) - 1) / sq_quads(2));
      k_searchq<2><<<g2, sq_warps(2) * 32, searchq_smem_bytes(2), s>>>(P);
    } else {
      const int w2 = search_warps(2);
      const int g2 = (int)std::min<u64>((u64)c->search_blocks2, ((u64)c->n_rows + w2 - 1) / w2);
      k_search<2><<<g2, w2 * 32, search_smem_total(2), s>>>(P);
    }
    c->launches += 2;
  }
  CK(cudaEventRecord(c->ev[EV_P2], s));
  {  // (5a) count symbols per macroblock, scan per image
    c->tok_grid = dim3((c->max_mb + TOK_WARPS - 1) / TOK_WARPS, ni);
    k_tokenize<0><<<c->tok_grid, TOK_WARPS * 32, 0, s>>>(P);
    k_tokscan<<<ni, 256, 0, s>>>(P);
    c->launches += 2;
  }
  int rc = lane_launch_b(c);
  if (rc != ZW_OK) return rc;
  c->state = LANE_IN_FLIGHT;
  return ZW_OK;
}

// Wait for the per-image results of a launched batch; re-run phase B with larger arenas in the (rare) case the
// symbol estimate was too small.  Fills the per-stage device times.
static int lane_sync_sizes(Lane* c) {
  if (c->n_valid == 0) { c->state = LANE_SIZED; return ZW_OK; }
  CK(cudaEventSynchronize(c->ev[EV_SIZES]));
  const ChunkTotals* T = c->h_tot.as<ChunkTotals>();
  while (T->overflow) {
    if (c->reruns++ > 2) return g_last_error = ZW_ERR_OUTPUT_TOO_SMALL;
    c->tok_per_mb = std::max(c->tok_per_mb * 1.25, 1.15 * (double)T->tok_tokens / (double)c->n_mb);
    c->cap_t = (u64)(1.15 * (double)T->tok_tokens) + 64;
    int rc = lane_launch_b(c);
    if (rc != ZW_OK) return rc;
    CK(cudaEventSynchronize(c->ev[EV_SIZES]));
  }
  CK(cudaGetLastError());
  zw_timing Tm = zw_timing();  // per-call device times; keeps the staged chunk's H2D figures
  Tm.h2d_bytes = c->last.h2d_bytes; Tm.pixels = c->last.pixels;
  auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]); return ms; };
  Tm.yuv_ms = el(EV_YUV0, EV_YUV1); Tm.analysis_ms = el(EV_YUV1, EV_AN1); Tm.chroma1_ms = el(EV_AN1, EV_C1);
  Tm.pass1_ms = el(EV_P1S, EV_P1); Tm.stats_ms = el(EV_F1S, EV_ST1);  // k_finish1 + statistics + probabilities
  Tm.chroma2_ms = el(EV_ST1, EV_C2); Tm.pass2_ms = el(EV_C2, EV_P2); Tm.token_ms = el(EV_P2, EV_TOK);
  Tm.boolcode_ms = el(EV_TOK, EV_BC); Tm.assemble_ms = el(EV_BC, EV_END); Tm.device_total_ms = el(EV_START, EV_END);
  Tm.h2d_ms = el(EV_H2D0, EV_H2D1);
  Tm.kernel_launches = c->launches;
  const ImageState* st = c->h_st.as<ImageState>();
  for (u32 i = 0; i < c->n_valid; i++) Tm.symbols += (u64)st[i].hdr_tokens + st[i].tok_tokens;
#ifdef ZW_WAIT_STATS
  {
    unsigned long long w[8];
    cudaMemcpy(w, c->d_ticket.p, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "wait stats: pass1 wait %.1f%% of row time, pass2 wait %.1f%%\n", 100.0 * (double)w[3] / (double)(w[5] ? w[5] : 1),
            100.0 * (double)w[4] / (double)(w[6] ? w[6] : 1));
  }
#endif
  c->last = Tm;
  c->state = LANE_SIZED;
  return ZW_OK;
}

// One D2H of the finished files into the lane's pinned arena + the per-image view (offset, length, status).
static int lane_fetch(Lane* c, int container) {
  const u32 ni = c->n_valid;
  cudaStream_t s = c->stream;
  const ImageState* st = c->h_st.as<ImageState>();
  const u64* off = c->h_outoff.as<u64>();
  if (ni) {
    const size_t total = (size_t)off[ni];
    CK(c->h_arena.reserve(total + 64));
    CK(cudaEventRecord(c->ev[EV_D2H0], s));
    CK(cudaMemcpyAsync(c->h_arena.p, c->d_out.p, total, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(c->ev[EV_D2H1], s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[EV_D2H0], c->ev[EV_D2H1]);
    c->last.d2h_ms = ms;
    c->last.d2h_bytes = total + ni * sizeof(ImageState) + ((size_t)ni + 1) * 8 + sizeof(ChunkTotals);
  }
  const size_t n = c->n_in;
  c->v_off.assign(n, 0); c->v_len.assign(n, 0); c->v_status.assign(n, 0);
  for (size_t i = 0; i < n; i++) {
    if (c->slot_of[i] < 0) { c->v_status[i] = c->img_status[i]; continue; }
    const u32 k = (u32)c->slot_of[i];
    // lossy + alpha files need VP8X + ALPH (api.rs:1330-1394), which is not built: refuse rather than
    // emit a simple container the reference would not produce
    if (container && (c->img[k].bpp == 2 || c->img[k].bpp == 4)) { c->v_status[i] = ZW_ERR_INVALID_PARAM; continue; }
    if (st[k].status != 0) { c->v_status[i] = (int)st[k].status; continue; }
    c->v_off[i] = off[k] + (container ? 0 : 20);
    c->v_len[i] = container ? st[k].file_bytes : st[k].vp8_bytes;
  }
  c->container = container;
  c->state = LANE_DONE;
  return ZW_OK;
}

// Copy a finished lane's files into caller-style output slots (malloc'ed when data == NULL).
static void lane_emit(const Lane* c, zw_output* outs) {
  const u8* arena = c->h_arena.as<u8>();
  for (size_t i = 0; i < c->n_in; i++) {
    zw_output& o = outs[i];
    o.len = 0;
    o.status = c->v_status[i];
    if (o.status != ZW_OK) continue;
    const size_t need = c->v_len[i];
    if (o.data == nullptr) {
      o.data = (uint8_t*)malloc(need ? need : 1);
      o.cap = need;
      if (!o.data) { o.status = ZW_ERR_CUDA + (int)cudaErrorMemoryAllocation; continue; }
    } else if (o.cap < need) { o.status = ZW_ERR_OUTPUT_TOO_SMALL; o.len = need; continue; }
    memcpy(o.data, arena + c->v_off[i], need);
    o.len = need;
  }
}

// Integer issue peak (measurement only, SURVEY.md 8(d) "INT peak: measure, don't assume"): eight
// independent chains per thread, alternating the fma pipe (IMAD) and the alu pipe (LOP3 / IADD3),
// 48 integer instructions per thread and outer iteration, full occupancy.
__global__ void __launch_bounds__(256) k_intpeak(u32* out, int iters) {
  u32 a = threadIdx.x, b = blockIdx.x + 1, c = 3, d = 5, e = 7, f = 11, g = 13, h = 17;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      a = a * 5 + b;  c = c * 3 + d;  e = e * 7 + f;  g = g * 9 + h;    // 4 IMAD
      b = (b ^ c) + 1; d = (d ^ e) + 3; f = (f ^ g) + 5; h = (h ^ a) + 7; // 4 (LOP3, IADD) pairs; may fuse
    }
  }
  if ((a ^ b ^ c ^ d ^ e ^ f ^ g ^ h) == 0x12345678u) out[0] = a;
}

static void add_timing(zw_timing& a, const zw_timing& L) {
  a.h2d_ms += L.h2d_ms; a.yuv_ms += L.yuv_ms; a.analysis_ms += L.analysis_ms; a.pass1_ms += L.pass1_ms;
  a.stats_ms += L.stats_ms; a.pass2_ms += L.pass2_ms; a.token_ms += L.token_ms; a.boolcode_ms += L.boolcode_ms;
  a.assemble_ms += L.assemble_ms; a.d2h_ms += L.d2h_ms; a.chroma1_ms += L.chroma1_ms; a.chroma2_ms += L.chroma2_ms;
  a.device_total_ms += L.device_total_ms;
  a.kernel_launches += L.kernel_launches; a.h2d_bytes += L.h2d_bytes; a.d2h_bytes += L.d2h_bytes; a.pixels += L.pixels;
  a.symbols += L.symbols;
}

static int check_params(int quality, int method) {
  return (quality < 0 || quality > 100 || method < 0) ? ZW_ERR_INVALID_PARAM : ZW_OK;
}

// Hand one batch (one chunk) to lane `k`: stage + launch everything, asynchronous.
static int ctx_submit_lane(zw_ctx* c, int k, const zw_image* imgs, size_t n, int quality, int method) {
  Lane* l = c->lanes[k];
  int rc = lane_stage(l, imgs, n, c->prev_copy);
  if (rc != ZW_OK) return rc;
  if (l->n_valid) c->prev_copy = l;
  rc = lane_launch(l, quality, std::min(method, 6) /* vp8.rs:1291 */, c->prev);
  if (rc != ZW_OK) { l->state = LANE_FREE; return rc; }
  if (l->n_valid) c->prev = l;
  c->dump_lane = k;
  return ZW_OK;
}

struct zw_multi {
  std::vector<zw_ctx*> ctx;
};

extern "C" {

const char* zw_version(void) { return "zenwebp_b200 0.2 (CUDA, sm_100a)"; }
int zw_last_error(void) { return g_last_error; }
void zw_free(void* p) { free(p); }
size_t zw_max_output_size(uint32_t w, uint32_t h) {
  const size_t mbw = (w + 15) / 16, mbh = (h + 15) / 16;
  return 64 + 16384 + mbw * mbh * (7700 * 7 / 8 + 200);
}
const char* zw_strerror(int code) {
  switch (code) {
    case ZW_OK: return "ok";
    case ZW_ERR_INVALID_DIMENSIONS: return "invalid dimensions";
    case ZW_ERR_INVALID_BUFFER_SIZE: return "invalid buffer size";
    case ZW_ERR_INVALID_PARAM: return "invalid parameter";
    case ZW_ERR_OUTPUT_TOO_SMALL: return "output buffer too small";
    case ZW_ERR_PARTITION_TOO_LARGE: return "first partition exceeds the 19-bit size field";
    case ZW_ERR_NOT_STAGED: return "no staged batch / no such ticket";
    case ZW_ERR_BUSY: return "every pipeline slot of the context holds a batch (wait + release one first)";
    case ZW_ERR_TOO_LARGE: return "batch exceeds the per-chunk device budget (split it, or use zw_encode_*_batch)";
    default: break;
  }
  if (code >= ZW_ERR_CUDA) return cudaGetErrorString((cudaError_t)(code - ZW_ERR_CUDA));
  return "unknown error";
}

zw_ctx* zw_create(int device, const zw_limits* limits) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { g_last_error = ZW_ERR_CUDA + (int)(e != cudaSuccess ? e : cudaErrorNoDevice); return nullptr; }
  if (device < 0 || device >= ndev) { g_last_error = ZW_ERR_INVALID_PARAM; return nullptr; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_last_error = ZW_ERR_CUDA + (int)e; return nullptr; }
  zw_ctx* c = new zw_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_last_error = ZW_ERR_CUDA + (int)e; delete c; return nullptr; }
  c->sm_count = prop.multiProcessorCount;
  c->budget = (limits && limits->max_device_bytes) ? limits->max_device_bytes : ((size_t)32 << 30);
  int warps_hint = limits ? limits->persistent_warps_per_sm : 0;
  if (const char* env = getenv("ZW_WARPS_PER_SM")) warps_hint = atoi(env);
  int depth = limits ? limits->reserved[0] : 0;  // reserved[0]: pipeline depth = batches in flight (0 = default 3)
  if (const char* env = getenv("ZW_LANES")) depth = atoi(env);
  if (depth <= 0) depth = 3;
  depth = std::min(depth, 8);
  // constant tables
  std::vector<SegParams> segtab(128);
  for (int i = 0; i < 128; i++) segtab[i] = make_segparams(i);
  std::vector<u8> lut(128 * 255);
  for (int b = 0; b < 128; b++)
    for (int a = -127; a <= 127; a++) lut[b * 255 + (a + 127)] = (u8)compute_segment_quant(b, a, 50);
  const TokenTables tt = make_token_tables();
  if ((e = c->d_segtab.reserve(segtab.size() * sizeof(SegParams))) != cudaSuccess || (e = c->d_lut.reserve(lut.size())) != cudaSuccess ||
      (e = cudaMemcpy(c->d_segtab.p, segtab.data(), segtab.size() * sizeof(SegParams), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(c->d_lut.p, lut.data(), lut.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpyToSymbol(c_tok, &tt, sizeof(tt))) != cudaSuccess) {
    g_last_error = ZW_ERR_CUDA + (int)e; zw_destroy(c); return nullptr;
  }
  for (int k = 0; k < depth; k++) {
    Lane* l = lane_create(c, warps_hint);
    if (!l) {
      e = cudaGetLastError();
      g_last_error = ZW_ERR_CUDA + (int)(e != cudaSuccess ? e : cudaErrorMemoryAllocation); zw_destroy(c); return nullptr;
    }
    c->lanes.push_back(l);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_error = ZW_ERR_CUDA + (int)e; zw_destroy(c); return nullptr; }
  g_last_error = 0;
  return c;
}

void zw_destroy(zw_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (Lane* l : c->lanes) lane_destroy(l);
  c->dec.release();
  c->ll.release();
  c->d_segtab.release(); c->d_lut.release();
  delete c;
}

// ---- streaming entry points: one context keeps `depth` batches in flight ---------------------------------
int zw_submit(zw_ctx* c, const zw_image* imgs, size_t n, int quality, int method, int* ticket) {
  if (!c || (!imgs && n) || !ticket) return g_last_error = ZW_ERR_INVALID_PARAM;
  if (check_params(quality, method) != ZW_OK) return g_last_error = ZW_ERR_INVALID_PARAM;
  if (n > 65535) return g_last_error = ZW_ERR_TOO_LARGE;  // grid.y limit of the per-image kernels
  size_t bytes = 0;
  for (size_t i = 0; i < n; i++)
    if (validate_image(imgs[i]) == ZW_OK) bytes += image_footprint(imgs[i].width, imgs[i].height, color_bpp(imgs[i].color));
  if (n > 1 && bytes > c->budget) return g_last_error = ZW_ERR_TOO_LARGE;
  CK(cudaSetDevice(c->device));
  if (c->staged) {  // lane 0 doubles as the split API's lane: a batch left staged there is dropped, not a taken slot
    c->lanes[0]->state = LANE_FREE;
    c->staged = false; c->encoded = false;
  }
  int k = -1;
  for (size_t i = 0; i < c->lanes.size(); i++)
    if (c->lanes[i]->state == LANE_FREE) { k = (int)i; break; }
  if (k < 0) return g_last_error = ZW_ERR_BUSY;
  int rc = ctx_submit_lane(c, k, imgs, n, quality, method);
  if (rc != ZW_OK) return g_last_error = rc;
  *ticket = k;
  return g_last_error = ZW_OK;
}

int zw_wait(zw_ctx* c, int ticket, int container, zw_batch_view* view, zw_timing* timing) {
  if (!c || ticket < 0 || ticket >= (int)c->lanes.size()) return g_last_error = ZW_ERR_INVALID_PARAM;
  Lane* l = c->lanes[ticket];
  if (l->state != LANE_IN_FLIGHT && l->state != LANE_SIZED && l->state != LANE_DONE) return g_last_error = ZW_ERR_NOT_STAGED;
  CK(cudaSetDevice(c->device));
  int rc;
  if (l->state == LANE_IN_FLIGHT && (rc = lane_sync_sizes(l)) != ZW_OK) return g_last_error = rc;
  if (l->state != LANE_DONE || l->container != (container != 0)) {
    if ((rc = lane_fetch(l, container != 0)) != ZW_OK) return g_last_error = rc;
  }
  if (view) {
    view->arena = l->h_arena.as<uint8_t>();
    view->n = l->n_in;
    view->offsets = l->v_off.data();
    view->lens = l->v_len.data();
    view->status = l->v_status.data();
  }
  if (timing) *timing = l->last;
  return g_last_error = ZW_OK;
}

int zw_release(zw_ctx* c, int ticket) {
  if (!c || ticket < 0 || ticket >= (int)c->lanes.size()) return g_last_error = ZW_ERR_INVALID_PARAM;
  Lane* l = c->lanes[ticket];
  if (l->state == LANE_IN_FLIGHT) {  // abandon: let the device finish before the buffers are reused
    cudaSetDevice(c->device);
    cudaStreamSynchronize(l->stream);
  }
  l->state = LANE_FREE;
  return g_last_error = ZW_OK;
}

// ---- split form on lane 0 (kernel-only timing; inputs stay resident) ---------------------------------------
int zw_stage_batch(zw_ctx* c, const zw_image* imgs, size_t n) {
  if (!c || (!imgs && n)) return g_last_error = ZW_ERR_INVALID_PARAM;
  if (n > 65535) return g_last_error = ZW_ERR_TOO_LARGE;
  CK(cudaSetDevice(c->device));
  c->staged = false; c->encoded = false;
  Lane* l = c->lanes[0];
  if (l->state == LANE_IN_FLIGHT) CK(cudaStreamSynchronize(l->stream));
  int rc = lane_stage(l, imgs, n, c->prev_copy);
  if (rc != ZW_OK) return g_last_error = rc;
  if (l->n_valid) c->prev_copy = l;
  // the caller may reuse or free its buffers as soon as this returns: wait for the copies
  if (l->n_valid) CK(cudaEventSynchronize(l->ev[EV_H2D1]));
  float ms = 0;
  if (l->n_valid && cudaEventElapsedTime(&ms, l->ev[EV_H2D0], l->ev[EV_H2D1]) == cudaSuccess) l->last.h2d_ms = ms;
  c->n_staged = n;
  c->staged = true;
  return g_last_error = ZW_OK;
}

int zw_encode_resident(zw_ctx* c, int quality, int method, zw_timing* timing) {
  if (!c) return g_last_error = ZW_ERR_INVALID_PARAM;
  if (!c->staged) return g_last_error = ZW_ERR_NOT_STAGED;
  if (check_params(quality, method) != ZW_OK) return g_last_error = ZW_ERR_INVALID_PARAM;
  CK(cudaSetDevice(c->device));
  c->encoded = false;
  Lane* l = c->lanes[0];
  const zw_timing keep = l->last;
  int rc = lane_launch(l, quality, std::min(method, 6), c->prev);
  if (rc != ZW_OK) return g_last_error = rc;
  if (l->n_valid) c->prev = l;
  c->dump_lane = 0;
  if ((rc = lane_sync_sizes(l)) != ZW_OK) return g_last_error = rc;
  l->last.h2d_ms = keep.h2d_ms;
  c->last = l->last;
  if (timing) *timing = c->last;
  c->encoded = true;
  return g_last_error = ZW_OK;
}

int zw_download(zw_ctx* c, zw_output* outs, size_t n, int container, zw_timing* timing) {
  if (!c || (!outs && n)) return g_last_error = ZW_ERR_INVALID_PARAM;
  if (!c->staged || !c->encoded) return g_last_error = ZW_ERR_NOT_STAGED;
  if (n != c->n_staged) return g_last_error = ZW_ERR_INVALID_PARAM;
  CK(cudaSetDevice(c->device));
  Lane* l = c->lanes[0];
  int rc = lane_fetch(l, container != 0);
  if (rc != ZW_OK) return g_last_error = rc;
  lane_emit(l, outs);
  l->state = LANE_SIZED;  // stays staged + encoded: download / re-encode again at will
  c->last = l->last;
  if (timing) *timing = c->last;
  return g_last_error = ZW_OK;
}

// ---- batch entry points: chunked by the device budget, chunks pipelined over the context's lanes ----------
static int encode_batch(zw_ctx* c, const zw_image* imgs, size_t n, int quality, int method, zw_output* outs, zw_timing* timing,
                        int container) {
  if (!c || (!imgs && n) || (!outs && n)) return g_last_error = ZW_ERR_INVALID_PARAM;
  if (check_params(quality, method) != ZW_OK) return g_last_error = ZW_ERR_INVALID_PARAM;
  CK(cudaSetDevice(c->device));
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t i = 0; i < n; i++) { outs[i].len = 0; outs[i].status = ZW_ERR_NOT_STAGED; }  // overwritten per chunk; what a failed call leaves behind
  for (Lane* l : c->lanes) {  // a batch call owns the whole context
    if (l->state == LANE_IN_FLIGHT) CK(cudaStreamSynchronize(l->stream));
    l->state = LANE_FREE;
  }
  c->staged = false; c->encoded = false;
  // Chunks: bounded by the device budget.  A large batch is cut into up to THREE chunks of growing size (1 : 3 : 9):
  // only the first, small chunk's H2D copy is exposed, every later copy runs under the kernels of the chunk before it
  // (the link moves pixels ~3.7x faster than the kernels consume them), and most images still go through the kernels
  // in one large chunk.  Every extra chunk pays the per-chunk latency floor once more -- above all the pass-1 chroma
  // chains (one warp per image, ~2.4 us per macroblock: 3.7 ms for 768x512 images, 20 ms for 1920x1080) and about as
  // much again in wavefront ramps and tails, which also grow with the image size -- so the number of chunks is the one
  // with the smallest estimate of exposed copy + extra floors.  Measured, blocking call: 1024 x 768x512 m4 one chunk
  // 101.9 ms, 1:3 98.9 ms, 1:3:9 ~118 ms; 256 x 1920x1080 m6 one chunk 200.7 ms, 1:3 222.5 ms, 1:3:9 270 ms;
  // 8192 x 256x256 m0: 1:3:9 82.8 ms.  ZW_SPLIT=n forces n equal chunks (1 = no split).
  u64 total_px = 0, total_bytes = 0, max_mb = 0;
  for (size_t i = 0; i < n; i++) {
    total_px += (u64)imgs[i].width * imgs[i].height;
    total_bytes += imgs[i].len;
    max_mb = std::max<u64>(max_mb, (u64)((imgs[i].width + 15) / 16) * ((imgs[i].height + 15) / 16));
  }
  int split = 0;
  if (const char* env = getenv("ZW_SPLIT")) split = std::max(0, atoi(env));
  if (c->lanes.size() < 2) split = 1;
  std::vector<u64> px_targets;  // pixel budget of chunk 0, 1, ...; the last entry repeats
  if (split >= 1) px_targets.push_back((total_px + split - 1) / split);
  else {
    const double h2d_ms = (double)total_bytes / 50e6, floor_ms = (double)max_mb * 4.8e-3 + 2.0;
    const double est1 = h2d_ms, est2 = h2d_ms / 4 + floor_ms, est3 = h2d_ms / 13 + 2 * floor_ms;
    if (total_px >= (96ull << 20) && est3 <= est2 && est3 < est1) px_targets = {total_px / 13, 3 * total_px / 13, total_px};
    else if (total_px >= (24ull << 20) && est2 < est1) px_targets = {total_px / 4, total_px};
    else px_targets.push_back(total_px + 1);
  }
  struct Pending { int lane; size_t i0, i1; };
  std::vector<Pending> q;
  zw_timing acc = zw_timing();
  int rc = ZW_OK;
  auto drain_one = [&]() -> int {
    const Pending p = q.front();
    q.erase(q.begin());
    Lane* l = c->lanes[p.lane];
    int r = lane_sync_sizes(l);
    if (r == ZW_OK) r = lane_fetch(l, container);
    if (r == ZW_OK) { lane_emit(l, outs + p.i0); add_timing(acc, l->last); }
    else for (size_t i = p.i0; i < p.i1; i++) outs[i].status = r;
    l->state = LANE_FREE;
    return r;
  };
  size_t i0 = 0, chunk_no = 0;
  while (i0 < n && rc == ZW_OK) {
    size_t i1 = i0, bytes = 0;
    u64 px = 0;
    const u64 px_target = px_targets[std::min(chunk_no, px_targets.size() - 1)];
    chunk_no++;
    while (i1 < n && (i1 - i0) < 32768) {
      const bool ok = validate_image(imgs[i1]) == ZW_OK;
      const size_t f = ok ? image_footprint(imgs[i1].width, imgs[i1].height, color_bpp(imgs[i1].color)) : 0;
      if (i1 > i0 && (bytes + f > c->budget || px >= px_target)) break;
      bytes += f; px += ok ? (u64)imgs[i1].width * imgs[i1].height : 0; i1++;
    }
    int k = -1;
    for (;;) {
      for (size_t j = 0; j < c->lanes.size(); j++)
        if (c->lanes[j]->state == LANE_FREE) { k = (int)j; break; }
      if (k >= 0 || q.empty()) break;
      if ((rc = drain_one()) != ZW_OK) break;
    }
    if (rc != ZW_OK || k < 0) break;
    rc = ctx_submit_lane(c, k, imgs + i0, i1 - i0, quality, method);
    if (rc != ZW_OK) { for (size_t i = i0; i < i1; i++) outs[i].status = rc; break; }
    q.push_back({k, i0, i1});
    i0 = i1;
  }
  while (!q.empty()) { const int r = drain_one(); if (rc == ZW_OK) rc = r; }
  if (rc != ZW_OK) {  // chunks that never ran carry the failing code
    for (size_t i = i0; i < n; i++) if (outs[i].status == ZW_ERR_NOT_STAGED) outs[i].status = rc;
    return g_last_error = rc;
  }
  acc.wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  c->last = acc;
  if (timing) *timing = acc;
  return g_last_error = ZW_OK;
}

int zw_encode_vp8_batch(zw_ctx* c, const zw_image* imgs, size_t n, int quality, int method, zw_output* outs, zw_timing* timing) {
  return encode_batch(c, imgs, n, quality, method, outs, timing, 0);
}
int zw_encode_webp_batch(zw_ctx* c, const zw_image* imgs, size_t n, int quality, int method, zw_output* outs, zw_timing* timing) {
  return encode_batch(c, imgs, n, quality, method, outs, timing, 1);
}

// ---- one batch over several GPUs: contiguous slices of ceil(n / G) images, one host thread + context per GPU,
//      no collective; results land in outs in image order --------------------------------------
zw_multi* zw_multi_create(const int* devices, int n_devices, const zw_limits* limits) {
  if (!devices || n_devices <= 0) { g_last_error = ZW_ERR_INVALID_PARAM; return nullptr; }
  zw_multi* m = new zw_multi();
  for (int i = 0; i < n_devices; i++) {
    zw_ctx* c = zw_create(devices[i], limits);
    if (!c) { const int err = g_last_error; zw_multi_destroy(m); g_last_error = err; return nullptr; }
    m->ctx.push_back(c);
  }
  return m;
}
void zw_multi_destroy(zw_multi* m) {
  if (!m) return;
  for (zw_ctx* c : m->ctx) zw_destroy(c);
  delete m;
}
int zw_multi_device_count(const zw_multi* m) { return m ? (int)m->ctx.size() : 0; }
int zw_multi_encode(zw_multi* m, const zw_image* imgs, size_t n, int quality, int method, int container, zw_output* outs,
                    zw_timing* per_device /* [device_count] or NULL */) {
  if (!m || (!imgs && n) || (!outs && n)) return g_last_error = ZW_ERR_INVALID_PARAM;
  if (check_params(quality, method) != ZW_OK) return g_last_error = ZW_ERR_INVALID_PARAM;
  const size_t G = m->ctx.size(), per = (n + G - 1) / G;
  std::vector<int> rcs(G, ZW_OK);
  std::vector<std::thread> ts;
  for (size_t g = 0; g < G; g++) {
    const size_t b = std::min(n, g * per), e = std::min(n, b + per);
    if (per_device) per_device[g] = zw_timing();
    if (e == b) continue;
    ts.emplace_back([=, &rcs]() {
      rcs[g] = encode_batch(m->ctx[g], imgs + b, e - b, quality, method, outs + b, per_device ? per_device + g : nullptr, container);
    });
  }
  for (auto& t : ts) t.join();
  for (size_t g = 0; g < G; g++) if (rcs[g] != ZW_OK) return g_last_error = rcs[g];
  return g_last_error = ZW_OK;
}

int zw_measure_int_peak(zw_ctx* c, double* int_instr_per_s) {
  if (!c || !int_instr_per_s) return g_last_error = ZW_ERR_INVALID_PARAM;
  CK(cudaSetDevice(c->device));
  Lane* l = c->lanes[0];
  CK(l->d_ticket.reserve(64));
  const int iters = 20000, grid = c->sm_count * 8;
  // instructions per thread and outer iteration, read off the SASS of k_intpeak (cuobjdump -sass):
  // 16 IMAD (fma pipe) + 16 LOP3 + 16 IADD3/VIADD (alu pipe) = 48; loop overhead not counted
  const double per_iter = 48.0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_intpeak<<<grid, 256, 0, l->stream>>>(l->d_ticket.as<u32>() + 15, 100);  // warm-up
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0, l->stream));
    k_intpeak<<<grid, 256, 0, l->stream>>>(l->d_ticket.as<u32>() + 15, iters);
    CK(cudaEventRecord(e1, l->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  CK(cudaGetLastError());
  *int_instr_per_s = (double)grid * 256.0 * (double)iters * per_iter / ((double)best * 1e-3);
  return g_last_error = ZW_OK;
}

int zw_dump_stage(zw_ctx* ctx, size_t index, const char* stage, void* dst, size_t cap, size_t* len) {
  if (!ctx || !stage || !len) return g_last_error = ZW_ERR_INVALID_PARAM;
  if (ctx->dump_lane < 0) return g_last_error = ZW_ERR_NOT_STAGED;
  Lane* c = ctx->lanes[ctx->dump_lane];
  if (index >= c->n_in || c->slot_of[index] < 0 || c->state == LANE_IN_FLIGHT) return g_last_error = ZW_ERR_NOT_STAGED;
  CK(cudaSetDevice(ctx->device));
  const u32 k = (u32)c->slot_of[index];
  const ImageDesc& d = c->img[k];
  const size_t nmb = (size_t)d.mbw * d.mbh, ysz = nmb * 256, csz = nmb * 64;
  const std::string s(stage);
  const void* src = nullptr;
  size_t bytes = 0;
  std::vector<u8> tmp;
  ImageState st;
  ImageLayout lay;
  CK(cudaMemcpy(&st, c->d_st.as<ImageState>() + k, sizeof(ImageState), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&lay, c->d_lay.as<ImageLayout>() + k, sizeof(ImageLayout), cudaMemcpyDeviceToHost));
  if (s == "YUV_Y") { src = c->d_planes.as<u8>() + d.y_off; bytes = ysz; }
  else if (s == "YUV_U") { src = c->d_planes.as<u8>() + d.y_off + ysz; bytes = csz; }
  else if (s == "YUV_V") { src = c->d_planes.as<u8>() + d.y_off + ysz + csz; bytes = csz; }
  else if (s == "ALPHA") { src = c->d_alpha.as<u8>() + d.mb_off; bytes = nmb; }
  else if (s == "ALPHA_HIST") { src = c->d_alpha_hist.as<u8>() + (size_t)k * 1024; bytes = 1024; }
  else if (s == "SEG_MAP256") { src = c->d_map256.as<u8>() + (size_t)k * 256; bytes = 256; }
  else if (s == "SEG_MAP") { src = c->d_segmap.as<u8>() + d.mb_off; bytes = nmb; }
  else if (s == "P1MB") { src = c->d_rec1.as<MbRecord>() + d.mb_off; bytes = nmb * sizeof(MbRecord); }
  else if (s == "P2MB") { src = c->d_rec2.as<MbRecord>() + d.mb_off; bytes = nmb * sizeof(MbRecord); }
  else if (s == "STATS") { src = c->d_stats.as<u8>() + (size_t)k * 1056 * 4; bytes = 1056 * 4; }
  else if (s == "PROBS") { src = c->d_probs.as<u8>() + (size_t)k * 1056; bytes = 1056; }
  else if (s == "LCOST") { src = c->d_lcost.as<u8>() + (size_t)k * 6528 * 2; bytes = 6528 * 2; }
  else if (s == "PART0") { src = c->d_part.as<u8>() + lay.part_off; bytes = st.part0_bytes; }
  else if (s == "PART1") { src = c->d_part.as<u8>() + lay.part_off + lay.p0_cap; bytes = st.part1_bytes; }
  else if (s == "HDR_TOKENS") { src = c->d_htok.as<Token>() + lay.hdr_off; bytes = (size_t)st.hdr_tokens * 2; }
  else if (s == "TOK_TOKENS") { src = c->d_ttok.as<Token>() + lay.tok_off; bytes = (size_t)st.tok_tokens * 2; }
  else if (s == "VP8") { src = c->d_out.as<u8>() + lay.out_off + 20; bytes = st.vp8_bytes; }
  else if (s == "WEBP") { src = c->d_out.as<u8>() + lay.out_off; bytes = st.file_bytes; }
  else {
    // small scalar stages served from the host copy of ImageState
    if (s == "SEG_QIDX") tmp.assign(st.seg_qidx, st.seg_qidx + 4);
    else if (s == "SEG_TREE_PROBS") tmp.assign(st.tree_probs, st.tree_probs + 3);
    else if (s == "SEG_UPDATE_MAP") tmp.assign(1, st.update_map);
    else if (s == "SEG_ENABLED") tmp.assign(1, st.seg_enabled);
    else if (s == "SEG_CENTERS") tmp.assign(st.centers, st.centers + 4);
    else if (s == "SEG_MID") { tmp.resize(4); memcpy(tmp.data(), &st.mid_alpha, 4); }
    else if (s == "SKIP_PROB") tmp.assign(1, st.skip_prob);
    else if (s == "PROBS_UPDATED") tmp.assign(1, st.probs_updated);
    else if (s == "BASE_QIDX") tmp.assign(1, (u8)c->base_qidx);
    else return g_last_error = ZW_ERR_INVALID_PARAM;
    *len = tmp.size();
    if (cap < tmp.size()) return g_last_error = ZW_ERR_OUTPUT_TOO_SMALL;
    memcpy(dst, tmp.data(), tmp.size());
    return g_last_error = ZW_OK;
  }
  *len = bytes;
  if (cap < bytes) return g_last_error = ZW_ERR_OUTPUT_TOO_SMALL;
  if (bytes) CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return g_last_error = ZW_OK;
}

}  // extern "C"

#include "zw_dec_host.inc"
#include "zw_lossless_host.inc"
