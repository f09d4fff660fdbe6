// zw_lossless.cuh -- VP8L lossless encoder of the reference, batched, sm_100a.
//
// Reference (file:line under /root/reference):
//   encode_frame_lossless      src/encoder/api.rs:945-1167   (subtract-green + "top" predictor transforms, one Huffman
//                                                             code per channel, run-length back references of distance 1)
//   build_huffman_tree         src/encoder/api.rs:163-287    (+ std::collections::BinaryHeap for the tie order)
//   write_huffman_tree         src/encoder/api.rs:289-354,   write_single_entry_huffman_tree :152-161
//   length_to_symbol           src/encoder/api.rs:356-363,   count_run / write_run :366-421
//   BitWriter                  src/encoder/api.rs:109-149    (LSB-first bit packing)
//   encode_alpha_lossless      src/encoder/api.rs:1175-1222  (the alpha channel as an L8 image, implicit dimensions)
//
// The reference walks the pixels three times on one thread (transform, count, write).  Here every step is a
// data-parallel pass over TILES of LL_TILE consecutive pixels of an image (pixels are numbered row-major, runs may cross
// row ends exactly as in the reference's flat `chunks_exact(4)` walk):
//   k_ll_residual  residual ARGB of every pixel (closed form of the in-place transforms) + last run head of each tile
//   k_ll_carry     per image: run head carried into each tile (max-scan over tiles)
//   k_ll_tokens    per pixel: "emits a literal" / "emits a run of length r" (a run of N equal residuals is coded as groups of
//                  1 literal + up to 4096 repeats: pixel i of a run starting at s is a literal iff (i - s) % 4097 == 0 and
//                  carries the group's run token iff it is the group's last pixel) + the four histograms
//   k_ll_huffman   per image: the four Huffman codes (serial, <= 280 symbols: one lane per channel) + the header bits
//   k_ll_bits      bits per tile;  k_ll_scan: bit offset of every tile, size of every stream
//   k_ll_emit      every tile packs its codes in shared memory and ORs / stores the words into the (zeroed) output
// All of it is byte/integer work bound by HBM traffic (about 3 + 4 + 2 B/px written or read once, 4 + 2 B/px re-read twice).
// The per-pixel and Huffman functions are ZW_HD so that tests/hostcheck runs the same source on the CPU against the oracle.
#ifndef ZW_LOSSLESS_CUH
#define ZW_LOSSLESS_CUH
#include "zw_prims.cuh"

namespace zw {

constexpr int LL_THREADS = 256;
constexpr int LL_PPT = 4;                       // pixels per thread
constexpr int LL_TILE = LL_THREADS * LL_PPT;    // pixels per tile
constexpr int LL_HDR_WORDS = 288;               // header + four serialised trees: <= 43 + 67 + 4 * 2035 bits
constexpr int LL_TREE_WORDS = 66;               // one serialised tree: <= 1 + 4 + 57 + 12 + 280 * 7 bits
constexpr int LL_MAX_PIXEL_BITS = 4 * 15 + 15 + 10;  // literal (four codes) + run symbol + extra bits
constexpr u32 LL_RUN_GROUP = 4097;              // 1 literal + at most 4096 repeats (api.rs:368)

enum { LL_FLAG_PREDICTOR = 1, LL_FLAG_IMPLICIT_DIMS = 2, LL_FLAG_ALPHA_PLANE = 4 };

struct LlImage {
  u64 src_off;   // bytes into the source arena
  u64 px_off;    // pixels into the residual / token arenas
  u32 width, height, npx;
  u32 tile_off, n_tiles;
  u8 bpp;        // bytes per source pixel
  u8 color;      // coded colour type: 0 L8, 1 La8, 2 Rgb8, 3 Rgba8 (api.rs:83-92)
  u8 flags;      // LL_FLAG_*; ALPHA_PLANE: the coded L8 image is the last byte of every source pixel (api.rs:1200-1206)
  u8 pad;
};
struct LlState {
  u64 total_bits;  // header + pixel data, before the flush to a whole byte
  u32 hdr_bits;
  u32 bytes;       // stream length (api.rs:139-149 flush)
};
struct LlParams {
  const LlImage* img;
  LlState* st;
  const u8* src;
  u32* res;          // residual ARGB per pixel: byte 0 = R-G, 1 = G, 2 = B-G, 3 = A (the reference's pixel[0..3])
  u16* desc;         // per pixel: bit 0 literal, bits 1..13 run length carried (0 = none)
  u32* tile_last;    // per tile: 1 + pixel index of its last run head, 0 = none
  u32* tile_carry;   // per tile: 1 + pixel index of the last run head before the tile
  u32* tile_bits;
  u64* tile_bitoff;
  u32* hist;         // [n_img][4][280]
  u32* codes;        // [n_img][4][280]: length << 16 | code
  u32* hdr;          // [n_img][LL_HDR_WORDS]
  u8* out;           // zeroed before k_ll_emit
  const u64* out_off;  // [n_img] byte offset of the VP8L stream (multiple of 4); a container header goes 20 bytes before it
  u32 n_img, n_tiles;
  u32 container;     // 1: write the simple RIFF/WEBP/"VP8L" wrap (api.rs:1325-1329) in front of every stream
};

// ---- per-pixel arithmetic ---------------------------------------------------------------------------------------------
ZW_HD u32 ll_sub4(u32 a, u32 b) {  // per-byte wrapping subtraction
#if defined(__CUDA_ARCH__)
  return __vsub4(a, b);
#else
  return (((a | 0x80808080u) - (b & 0x7F7F7F7Fu)) ^ ((a ^ ~b) & 0x80808080u));
#endif
}
// The pixel after "expand to RGBA" + "subtract green" (api.rs:1002-1021).
ZW_HD u32 ll_sg_pixel(const u8* s, u32 i, u32 bpp, u32 color, u32 flags) {
  const u8* p = s + (u64)i * bpp;
  u32 r, g, b, a = 255;
  if (flags & LL_FLAG_ALPHA_PLANE) { r = g = b = p[bpp - 1]; }
  else if (color == 0) { r = g = b = p[0]; }
  else if (color == 1) { r = g = b = p[0]; a = p[1]; }
  else { r = p[0]; g = p[1]; b = p[2]; if (color == 3) a = p[3]; }
  return ((r - g) & 255u) | (g << 8) | (((b - g) & 255u) << 16) | (a << 24);
}
// Closed form of the in-place predictor transform (api.rs:1023-1037): rows >= 1 subtract the pixel above, row 0 the pixel
// to the left, pixel 0 opaque black -- each against the ORIGINAL (subtract-green) neighbour.
ZW_HD u32 ll_residual(const u8* s, u32 i, u32 x, u32 y, u32 w, u32 bpp, u32 color, u32 flags) {
  const u32 cur = ll_sg_pixel(s, i, bpp, color, flags);
  if (!(flags & LL_FLAG_PREDICTOR)) return cur;
  if (y > 0) return ll_sub4(cur, ll_sg_pixel(s, i - w, bpp, color, flags));
  if (x > 0) return ll_sub4(cur, ll_sg_pixel(s, i - 1, bpp, color, flags));
  return ll_sub4(cur, 0xFF000000u);
}
// What pixel i of a run that started at pixel s emits (count_run / write_run, api.rs:366-421).
ZW_HD u32 ll_token_desc(u32 i, u32 s, bool next_breaks) {
  const u32 k = (i - s) % LL_RUN_GROUP;
  u32 d = k == 0 ? 1u : 0u;
  if (k > 0 && (next_breaks || k == LL_RUN_GROUP - 1)) d |= k << 1;
  return d;
}
// Green-alphabet symbol + extra bits of a run (api.rs:356-363, :375-380).
ZW_HD void ll_run_symbol(u32 run, u32& symbol, u32& extra_bits, u32& extra_val) {
  if (run <= 4) { symbol = 256 + run - 1; extra_bits = 0; extra_val = 0; return; }
  const u32 len = run - 1;
  u32 hb = 0;
  while ((len >> (hb + 1)) != 0) hb++;
  const u32 second = (len >> (hb - 1)) & 1u;
  extra_bits = hb - 1;
  symbol = 256 + 2 * hb + second;
  extra_val = len & ((1u << extra_bits) - 1u);
}
ZW_HD bool ll_is_color(u32 color) { return color >= 2; }
ZW_HD bool ll_is_alpha(u32 color) { return color == 1 || color == 3; }

// ---- bit writer into pre-zeroed 32-bit words (LSB first, like api.rs:125-137) ---------------------------------------------
struct LlBits {
  u32* w;
  u32 pos;
  ZW_HD void put(u32 bits, u32 n) {  // n <= 16
    if (n == 0) return;
    const u32 k = pos >> 5, sh = pos & 31;
    w[k] |= bits << sh;
    if (sh + n > 32) w[k + 1] |= bits >> (32 - sh);
    pos += n;
  }
  ZW_HD void append(const u32* src, u32 nbits) {
    for (u32 i = 0; i < nbits; i += 16) {
      const u32 n = nbits - i < 16 ? nbits - i : 16;
      const u32 k = i >> 5, sh = i & 31;
      u32 v = src[k] >> sh;
      if (sh + n > 32) v |= src[k + 1] << (32 - sh);
      put(v & ((1u << n) - 1u), n);
    }
  }
};

// ---- Huffman construction ---------------------------------------------------------------------------------------------
// Scratch of one tree (shared memory on the device).
struct LlHuffScratch {
  u32 hfreq[280];       // heap: frequency
  u16 hidx[280];        // heap: node index (leaf < n, internal node n + k)
  u16 left[280], right[280];
  u8 depth[280];        // depth of internal node k
  u16 order[280];       // length-limiting branch only
};
// `impl Ord for Item` (api.rs:176-190) compares frequencies in reverse: a <= b  <=>  a.freq >= b.freq.
struct LlHeap {
  u32* f; u16* x; u32 len;
  ZW_HD void sift_down_range(u32 pos, u32 end) {
    const u32 ef = f[pos]; const u16 ex = x[pos];
    u32 child = 2 * pos + 1;
    while (child <= (end >= 2 ? end - 2 : 0)) {
      child += f[child] >= f[child + 1] ? 1 : 0;
      if (ef <= f[child]) { f[pos] = ef; x[pos] = ex; return; }
      f[pos] = f[child]; x[pos] = x[child]; pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1 && ef > f[child]) { f[pos] = f[child]; x[pos] = x[child]; pos = child; }
    f[pos] = ef; x[pos] = ex;
  }
  ZW_HD void rebuild() { for (u32 n = len / 2; n > 0;) { n--; sift_down_range(n, len); } }
  ZW_HD void pop(u32& of, u16& ox) {  // swap the last element into the root, sink it to the bottom, sift it up
    len--;
    of = f[len]; ox = x[len];
    if (len == 0) return;
    { const u32 tf = f[0]; const u16 tx = x[0]; f[0] = of; x[0] = ox; of = tf; ox = tx; }
    const u32 end = len;
    const u32 ef = f[0]; const u16 ex = x[0];
    u32 pos = 0, child = 1;
    while (child <= (end >= 2 ? end - 2 : 0)) {
      child += f[child] >= f[child + 1] ? 1 : 0;
      f[pos] = f[child]; x[pos] = x[child]; pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1) { f[pos] = f[child]; x[pos] = x[child]; pos = child; }
    while (pos > 0) {
      const u32 parent = (pos - 1) / 2;
      if (ef >= f[parent]) break;
      f[pos] = f[parent]; x[pos] = x[parent]; pos = parent;
    }
    f[pos] = ef; x[pos] = ex;
  }
};
ZW_HD u32 ll_rev16(u32 v) {
  v = ((v >> 1) & 0x5555u) | ((v & 0x5555u) << 1);
  v = ((v >> 2) & 0x3333u) | ((v & 0x3333u) << 2);
  v = ((v >> 4) & 0x0F0Fu) | ((v & 0x0F0Fu) << 4);
  return ((v >> 8) & 0x00FFu) | ((v & 0x00FFu) << 8);
}
// Symbols in ascending frequency, equal frequencies by index (what the length-limiting branch of build_huffman_tree
// walks, api.rs:248-260), as a rank sort shared by `nlanes` lanes: order[rank(i)] = i.  A photograph's residual
// histogram regularly needs it (rare symbols sit deeper than 15), and a serial insertion sort of 256 entries costs
// 2.4 M cycles on one lane -- more than the rest of the Huffman construction together.
ZW_HD void ll_rank_order(const u32* freq, u32 n, u16* order, u32 lane, u32 nlanes) {
  for (u32 i = lane; i < n; i += nlanes) {
    const u32 fi = freq[i];
    u32 r = 0;
    for (u32 j = 0; j < n; j++) r += (freq[j] < fi || (freq[j] == fi && j < i)) ? 1u : 0u;
    order[r] = (u16)i;
  }
}
// build_huffman_tree (api.rs:163-287).  false (lengths and codes zeroed) when at most one symbol is used.
// order_ready: S.order already holds ll_rank_order of `freq` (the kernel computes it with the whole warp).
ZW_HD bool ll_build_huffman(const u32* freq, u32 n, u8* lengths, u16* codes, u32 limit, LlHuffScratch& S, bool order_ready = false) {
  u32 used = 0;
  for (u32 i = 0; i < n; i++) { lengths[i] = 0; codes[i] = 0; used += freq[i] > 0 ? 1 : 0; }
  if (used <= 1) return false;
  LlHeap h;
  h.f = S.hfreq; h.x = S.hidx; h.len = 0;
  for (u32 i = 0; i < n; i++) if (freq[i] > 0) { h.f[h.len] = freq[i]; h.x[h.len] = (u16)i; h.len++; }
  h.rebuild();
  u32 n_int = 0;
  while (h.len > 1) {
    u32 f1; u16 i1;
    h.pop(f1, i1);
    S.left[n_int] = i1; S.right[n_int] = h.x[0];
    n_int++;
    h.f[0] += f1; h.x[0] = (u16)(n_int + n - 1);
    h.sift_down_range(0, h.len);
  }
  // depths: a parent is always created after its children, so one pass from the root (the last node) down
  S.depth[n_int - 1] = 0;
  u32 max_len = 0;
  for (u32 k = n_int; k > 0;) {
    k--;
    const u32 d = (u32)S.depth[k] + 1;
    const u32 c[2] = {S.left[k], S.right[k]};
    for (int j = 0; j < 2; j++) {
      if (c[j] < n) { lengths[c[j]] = (u8)d; max_len = d > max_len ? d : max_len; }
      else S.depth[c[j] - n] = (u8)d;
    }
  }
  if (max_len > limit) {
    u32 counts[16];
    for (int i = 0; i < 16; i++) counts[i] = 0;
    for (u32 i = 0; i < n; i++) counts[lengths[i] < limit ? lengths[i] : limit]++;
    u32 total = 0;
    for (u32 i = 1; i <= limit; i++) total += counts[i] << (limit - i);
    while (total > (1u << limit)) {
      u32 i = limit - 1;
      while (counts[i] == 0) i--;
      counts[i] -= 1; counts[limit] -= 1; counts[i + 1] += 2;
      total -= 1;
    }
    if (!order_ready) {  // ascending frequency, equal frequencies by index (insertion sort: stable)
      for (u32 i = 0; i < n; i++) {
        u32 j = i;
        while (j > 0 && freq[S.order[j - 1]] > freq[i]) { S.order[j] = S.order[j - 1]; j--; }
        S.order[j] = (u16)i;
      }
    }
    u32 len = limit;
    for (u32 k = 0; k < n; k++) {
      const u32 i = S.order[k];
      if (freq[i] > 0) {
        while (counts[len] == 0) len--;
        lengths[i] = (u8)len;
        counts[len]--;
      }
    }
  }
  // canonical codes in (length, index) order, bit-reversed (api.rs:268-282)
  u32 next[17];
  {
    u32 cnt[17];
    for (int i = 0; i < 17; i++) cnt[i] = 0;
    for (u32 i = 0; i < n; i++) cnt[lengths[i]]++;
    u32 code = 0;
    for (u32 l = 1; l <= limit; l++) { next[l] = code; code = (code + cnt[l]) << 1; }
  }
  for (u32 i = 0; i < n; i++) {
    const u32 l = lengths[i];
    if (l) { codes[i] = (u16)(ll_rev16(next[l] & 0xFFFFu) >> (16 - l)); next[l]++; }
  }
  return true;
}
ZW_HD void ll_write_single_entry_tree(LlBits& w, u32 symbol) {  // api.rs:152-161
  w.put(1, 2);
  if (symbol <= 1) { w.put(0, 1); w.put(symbol, 1); }
  else { w.put(1, 1); w.put(symbol, 8); }
}
// write_huffman_tree (api.rs:289-354): builds the code of one channel and serialises it.
ZW_HD void ll_write_huffman_tree(LlBits& w, const u32* freq, u32 n, u8* lengths, u16* codes, LlHuffScratch& S, bool order_ready = false) {
  if (!ll_build_huffman(freq, n, lengths, codes, 15, S, order_ready)) {
    u32 symbol = 0;
    for (u32 i = 0; i < n; i++) if (freq[i] > 0) { symbol = i; break; }
    ll_write_single_entry_tree(w, symbol & 255u);
    return;
  }
  u8 cl_len[16];
  u16 cl_code[16];
  u32 cl_freq[16];
  for (int i = 0; i < 16; i++) cl_freq[i] = 0;
  for (u32 i = 0; i < n; i++) cl_freq[lengths[i]]++;
  const bool single = !ll_build_huffman(cl_freq, 16, cl_len, cl_code, 7, S);
  const u8 order[19] = {17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  w.put(0, 1);
  w.put(19 - 4, 4);
  for (int k = 0; k < 19; k++) {
    const u32 i = order[k];
    if (i > 15 || cl_freq[i] == 0) w.put(0, 3);
    else if (single) w.put(1, 3);
    else w.put(cl_len[i], 3);
  }
  if (n == 256) { w.put(1, 1); w.put(3, 3); w.put(254, 8); }
  else w.put(0, 1);
  if (!single)
    for (u32 i = 0; i < n; i++) w.put(cl_code[lengths[i]], cl_len[lengths[i]]);
}
// Everything in front of the first Huffman tree (api.rs:974-1000).
ZW_HD void ll_write_prefix(LlBits& w, u32 width, u32 height, u32 color, u32 flags) {
  if (!(flags & LL_FLAG_IMPLICIT_DIMS)) {
    w.put(0x2f, 8);
    w.put(width - 1, 14);
    w.put(height - 1, 14);
    w.put(ll_is_alpha(color) ? 1 : 0, 1);
    w.put(0, 3);
  }
  w.put(5, 3);  // subtract green
  if (flags & LL_FLAG_PREDICTOR) {
    w.put(0x39, 6);  // predictor transform, 512-pixel blocks
    w.put(0, 1);     // no colour cache in the predictor image
    ll_write_single_entry_tree(w, 2);  // every block predicts from the pixel above
    for (int i = 0; i < 4; i++) ll_write_single_entry_tree(w, 0);
  }
  w.put(0, 1);  // transforms done
  w.put(0, 1);  // no colour cache
  w.put(0, 1);  // no meta-Huffman image
}
// Bits pixel `px` with token descriptor d adds to the stream, given the packed (length << 16 | code) tables.
template <class T>
ZW_HD u32 ll_pixel_bits(u32 px, u32 d, u32 color, T&& table /* (channel, symbol) -> length << 16 | code */) {
  u32 bits = 0;
  if (d & 1u) {
    bits += table(1, (px >> 8) & 255u) >> 16;
    if (ll_is_color(color)) bits += (table(0, px & 255u) >> 16) + (table(2, (px >> 16) & 255u) >> 16);
    if (ll_is_alpha(color)) bits += table(3, px >> 24) >> 16;
  }
  const u32 run = d >> 1;
  if (run) {
    u32 sym, eb, ev;
    ll_run_symbol(run, sym, eb, ev);
    bits += (table(1, sym) >> 16) + eb;
  }
  return bits;
}

#if defined(__CUDACC__)
// ---- kernels ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 ll_find_image(const LlImage* img, u32 n_img, u32 tile) {
  u32 lo = 0, hi = n_img - 1;
  while (lo < hi) {
    const u32 mid = (lo + hi + 1) >> 1;
    if (img[mid].tile_off <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ u32 ll_block_max_excl(u32 v, u32* sm /*[9]*/, u32& total) {  // exclusive max-scan over the CTA
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc = max(inc, t); }
  if (lane == 31) sm[warp] = inc;
  __syncthreads();
  u32 base = 0;
  for (u32 k = 0; k < warp; k++) base = max(base, sm[k]);
  u32 tot = 0;
  for (u32 k = 0; k < LL_THREADS / 32; k++) tot = max(tot, sm[k]);
  total = tot;
  u32 ex = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) ex = 0;
  __syncthreads();
  return max(base, ex);
}
__device__ __forceinline__ u32 ll_block_sum_excl(u32 v, u32* sm /*[9]*/, u32& total) {  // exclusive prefix sum over the CTA
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc += t; }
  if (lane == 31) sm[warp] = inc;
  __syncthreads();
  u32 base = 0, tot = 0;
  for (u32 k = 0; k < LL_THREADS / 32; k++) { if (k < warp) base += sm[k]; tot += sm[k]; }
  total = tot;
  __syncthreads();
  return base + inc - v;
}

// Residuals + the last run head of every tile.  One CTA per tile, four consecutive pixels per thread.
__global__ void __launch_bounds__(LL_THREADS) k_ll_residual(LlParams P) {
  __shared__ u32 s_last[LL_THREADS + 1];
  __shared__ u32 s_red[9];
  const u32 tile = blockIdx.x;
  const u32 ii = ll_find_image(P.img, P.n_img, tile);
  const LlImage im = P.img[ii];
  const u8* src = P.src + im.src_off;
  const u32 p0 = (tile - im.tile_off) * LL_TILE + threadIdx.x * LL_PPT;
  u32 r[LL_PPT];
  const u32 row_bytes = im.width * im.bpp;
  if (p0 + LL_PPT <= im.npx && p0 >= im.width && (row_bytes & 3u) == 0 && !(im.flags & LL_FLAG_ALPHA_PLANE) && (im.flags & LL_FLAG_PREDICTOR)) {
    // interior fast path: the four pixels and the four pixels above them are two runs of 4 * bpp bytes, both 4-byte
    // aligned (p0 is a multiple of 4, the source starts 16-byte aligned, row_bytes is a multiple of 4): word loads
    u32 cw[4], aw[4];
    const u32* cp = reinterpret_cast<const u32*>(src + (size_t)p0 * im.bpp);
    const u32* ap = reinterpret_cast<const u32*>(src + (size_t)(p0 - im.width) * im.bpp);
#pragma unroll
    for (int k = 0; k < 4; k++) { cw[k] = k < (int)im.bpp ? __ldg(cp + k) : 0u; aw[k] = k < (int)im.bpp ? __ldg(ap + k) : 0u; }
    auto byte_of = [](const u32* w, u32 b) { return (w[b >> 2] >> (8u * (b & 3u))) & 255u; };
#pragma unroll
    for (int j = 0; j < LL_PPT; j++) {
      u32 px[2];
#pragma unroll
      for (int t = 0; t < 2; t++) {
        const u32* w = t == 0 ? cw : aw;
        u32 rr, g, b, a = 255;
        if (im.bpp == 3) { rr = byte_of(w, 3 * j); g = byte_of(w, 3 * j + 1); b = byte_of(w, 3 * j + 2); }
        else if (im.bpp == 4) { rr = w[j] & 255u; g = (w[j] >> 8) & 255u; b = (w[j] >> 16) & 255u; a = w[j] >> 24; }
        else if (im.bpp == 1) { rr = g = b = byte_of(w, j); }
        else { rr = g = b = byte_of(w, 2 * j); a = byte_of(w, 2 * j + 1); }
        px[t] = ((rr - g) & 255u) | (g << 8) | (((b - g) & 255u) << 16) | (a << 24);
      }
      r[j] = ll_sub4(px[0], px[1]);
    }
  } else {
    u32 x = 0, y = 0;
    if (p0 < im.npx) { y = p0 / im.width; x = p0 - y * im.width; }
#pragma unroll
    for (int j = 0; j < LL_PPT; j++) {
      const u32 i = p0 + j;
      r[j] = 0;
      if (i < im.npx) {
        r[j] = ll_residual(src, i, x, y, im.width, im.bpp, im.color, im.flags);
        if (++x == im.width) { x = 0; y++; }
      }
    }
  }
  s_last[threadIdx.x + 1] = r[LL_PPT - 1];
  if (threadIdx.x == 0) {
    u32 prev = 0;
    if (p0 > 0 && p0 <= im.npx) {  // residual of the pixel in front of the tile
      const u32 i = p0 - 1, yy = i / im.width;
      prev = ll_residual(src, i, i - yy * im.width, yy, im.width, im.bpp, im.color, im.flags);
    }
    s_last[0] = prev;
  }
  __syncthreads();
  u32 prev = s_last[threadIdx.x];
  u32 last = 0;
  if (p0 < im.npx) {
    if (p0 + LL_PPT <= im.npx && ((im.px_off + p0) & 3) == 0)
      *reinterpret_cast<uint4*>(P.res + im.px_off + p0) = make_uint4(r[0], r[1], r[2], r[3]);
    else
      for (int j = 0; j < LL_PPT; j++) if (p0 + j < im.npx) P.res[im.px_off + p0 + j] = r[j];
#pragma unroll
    for (int j = 0; j < LL_PPT; j++) {
      const u32 i = p0 + j;
      if (i < im.npx && (i == 0 || r[j] != prev)) last = i + 1;
      prev = r[j];
    }
  }
  u32 tot;
  ll_block_max_excl(last, s_red, tot);
  if (threadIdx.x == 0) P.tile_last[tile] = tot;
}

// Run head carried into every tile: exclusive max-scan over the tiles of an image (one warp per image).
__global__ void __launch_bounds__(32) k_ll_carry(LlParams P) {
  const u32 ii = blockIdx.x, lane = threadIdx.x;
  const LlImage im = P.img[ii];
  u32 carry = 0;
  for (u32 t0 = 0; t0 < im.n_tiles; t0 += 32) {
    const u32 t = t0 + lane;
    const u32 v = t < im.n_tiles ? P.tile_last[im.tile_off + t] : 0;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc = max(inc, u); }
    u32 ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = 0;
    if (t < im.n_tiles) P.tile_carry[im.tile_off + t] = max(carry, ex);
    carry = max(carry, __shfl_sync(0xffffffffu, inc, 31));
  }
}

// Token descriptors + histograms.
__global__ void __launch_bounds__(LL_THREADS) k_ll_tokens(LlParams P) {
  __shared__ u32 s_hist[4 * 280];
  __shared__ u32 s_red[9];
  const u32 tile = blockIdx.x;
  const u32 ii = ll_find_image(P.img, P.n_img, tile);
  const LlImage im = P.img[ii];
  for (u32 k = threadIdx.x; k < 4 * 280; k += LL_THREADS) s_hist[k] = 0;
  const u32 p0 = (tile - im.tile_off) * LL_TILE + threadIdx.x * LL_PPT;
  const u32* res = P.res + im.px_off;
  u32 r[LL_PPT + 1];
  if (p0 + LL_PPT < im.npx) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(res + p0));
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = res[p0 + LL_PPT];
  } else {
#pragma unroll
    for (int j = 0; j <= LL_PPT; j++) r[j] = (p0 + j < im.npx) ? res[p0 + j] : 0;
  }
  const u32 prev = (p0 > 0 && p0 <= im.npx) ? res[p0 - 1] : 0;
  u32 last = 0;
  {
    u32 pv = prev;
#pragma unroll
    for (int j = 0; j < LL_PPT; j++) {
      const u32 i = p0 + j;
      if (i < im.npx && (i == 0 || r[j] != pv)) last = i + 1;
      pv = r[j];
    }
  }
  u32 tot;
  u32 head = ll_block_max_excl(last, s_red, tot);  // also orders the histogram zeroing before its use
  head = max(head, P.tile_carry[tile]);
  const bool color = ll_is_color(im.color), alpha = ll_is_alpha(im.color);
  u32 pv = prev;
  u32 d4[LL_PPT];
#pragma unroll
  for (int j = 0; j < LL_PPT; j++) {
    const u32 i = p0 + j;
    d4[j] = 0;
    if (i < im.npx) {
      if (i == 0 || r[j] != pv) head = i + 1;
      pv = r[j];
      const bool next_breaks = (i + 1 == im.npx) || r[j + 1] != r[j];
      const u32 d = ll_token_desc(i, head - 1, next_breaks);
      d4[j] = d;
      if (d & 1u) {
        atomicAdd(&s_hist[280 + ((r[j] >> 8) & 255u)], 1u);
        if (color) { atomicAdd(&s_hist[r[j] & 255u], 1u); atomicAdd(&s_hist[560 + ((r[j] >> 16) & 255u)], 1u); }
        if (alpha) atomicAdd(&s_hist[840 + (r[j] >> 24)], 1u);
      }
      if (d >> 1) {
        u32 sym, eb, ev;
        ll_run_symbol(d >> 1, sym, eb, ev);
        atomicAdd(&s_hist[280 + sym], 1u);
      }
    }
  }
  if (p0 < im.npx) {
    u16* dp = P.desc + im.px_off + p0;
    if (p0 + LL_PPT <= im.npx && ((im.px_off + p0) & 3) == 0)
      *reinterpret_cast<uint2*>(dp) = make_uint2(d4[0] | (d4[1] << 16), d4[2] | (d4[3] << 16));
    else
      for (int j = 0; j < LL_PPT; j++) if (p0 + j < im.npx) dp[j] = (u16)d4[j];
  }
  __syncthreads();
  u32* gh = P.hist + (size_t)ii * 4 * 280;
  for (u32 k = threadIdx.x; k < 4 * 280; k += LL_THREADS) { const u32 v = s_hist[k]; if (v) atomicAdd(&gh[k], v); }
}

// The Huffman codes of one image (warp c builds channel c: 0 red, 1 green + run lengths, 2 blue, 3 alpha) and its header.
struct LlHuffShared {
  LlHuffScratch scratch[4];
  u32 freq[4][280];
  u8 lengths[4][280];
  u16 codes[4][280];
  u32 tree[4][LL_TREE_WORDS];
  u32 tree_bits[4];
};
__global__ void __launch_bounds__(128) k_ll_huffman(LlParams P) {
  extern __shared__ __align__(16) unsigned char ll_smem[];
  LlHuffShared& S = *reinterpret_cast<LlHuffShared*>(ll_smem);
  const u32 ii = blockIdx.x, c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LlImage im = P.img[ii];
  const bool color = ll_is_color(im.color), alpha = ll_is_alpha(im.color);
  const bool built = c == 1 || (color && (c == 0 || c == 2)) || (alpha && c == 3);
  const u32 n = c == 1 ? 280 : 256;
  for (u32 k = lane; k < 280; k += 32) {
    S.freq[c][k] = P.hist[((size_t)ii * 4 + c) * 280 + k];
    S.lengths[c][k] = 0; S.codes[c][k] = 0;
  }
  for (u32 k = lane; k < LL_TREE_WORDS; k += 32) S.tree[c][k] = 0;
  __syncwarp();
  if (built) ll_rank_order(S.freq[c], n, S.scratch[c].order, lane, 32);
  __syncwarp();
  if (lane == 0) {
    LlBits w;
    w.w = S.tree[c]; w.pos = 0;
    if (built) ll_write_huffman_tree(w, S.freq[c], n, S.lengths[c], S.codes[c], S.scratch[c], true);
    else if (c == 3) ll_write_single_entry_tree(w, (im.flags & LL_FLAG_PREDICTOR) ? 0u : 255u);  // api.rs:1096-1100
    else ll_write_single_entry_tree(w, 0);
    S.tree_bits[c] = w.pos;
  }
  __syncwarp();
  for (u32 k = lane; k < 280; k += 32) P.codes[((size_t)ii * 4 + c) * 280 + k] = ((u32)S.lengths[c][k] << 16) | S.codes[c][k];
  __syncthreads();
  u32* hdr = P.hdr + (size_t)ii * LL_HDR_WORDS;
  for (u32 k = threadIdx.x; k < LL_HDR_WORDS; k += 128) hdr[k] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    LlBits w;
    w.w = hdr; w.pos = 0;
    ll_write_prefix(w, im.width, im.height, im.color, im.flags);
    w.append(S.tree[1], S.tree_bits[1]);  // green first (api.rs:1087-1101), then red, blue, alpha, distance
    w.append(S.tree[0], S.tree_bits[0]);
    w.append(S.tree[2], S.tree_bits[2]);
    w.append(S.tree[3], S.tree_bits[3]);
    ll_write_single_entry_tree(w, 1);
    P.st[ii].hdr_bits = w.pos;
  }
}

// The thread's four residuals and token descriptors: one 16-byte and one 8-byte load inside full, aligned groups.
__device__ __forceinline__ void ll_load4(const LlParams& P, const LlImage& im, u32 p0, u32* px, u32* d) {
  if (p0 + LL_PPT <= im.npx) {  // px_off and p0 are multiples of 4
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(P.res + im.px_off + p0));
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(P.desc + im.px_off + p0));
    px[0] = a.x; px[1] = a.y; px[2] = a.z; px[3] = a.w;
    d[0] = b.x & 0xFFFFu; d[1] = b.x >> 16; d[2] = b.y & 0xFFFFu; d[3] = b.y >> 16;
  } else {
#pragma unroll
    for (int j = 0; j < LL_PPT; j++) {
      const bool in = p0 + j < im.npx;
      px[j] = in ? P.res[im.px_off + p0 + j] : 0u;
      d[j] = in ? (u32)P.desc[im.px_off + p0 + j] : 0u;  // descriptor 0: the pixel emits nothing
    }
  }
}

// Bits of every tile.
__global__ void __launch_bounds__(LL_THREADS) k_ll_bits(LlParams P) {
  __shared__ u32 s_tab[4 * 280];
  __shared__ u32 s_red[9];
  const u32 tile = blockIdx.x;
  const u32 ii = ll_find_image(P.img, P.n_img, tile);
  const LlImage im = P.img[ii];
  for (u32 k = threadIdx.x; k < 4 * 280; k += LL_THREADS) s_tab[k] = P.codes[(size_t)ii * 4 * 280 + k];
  __syncthreads();
  const u32 p0 = (tile - im.tile_off) * LL_TILE + threadIdx.x * LL_PPT;
  u32 bits = 0;
  u32 px[LL_PPT], d[LL_PPT];
  ll_load4(P, im, p0, px, d);
#pragma unroll
  for (int j = 0; j < LL_PPT; j++)
    if (d[j]) bits += ll_pixel_bits(px[j], d[j], im.color, [&](u32 ch, u32 sym) { return s_tab[ch * 280 + sym]; });
  u32 tot;
  ll_block_sum_excl(bits, s_red, tot);
  if (threadIdx.x == 0) P.tile_bits[tile] = tot;
}

// Bit offset of every tile and the size of every stream (one warp per image).
__global__ void __launch_bounds__(32) k_ll_scan(LlParams P) {
  const u32 ii = blockIdx.x, lane = threadIdx.x;
  const LlImage im = P.img[ii];
  u64 run = P.st[ii].hdr_bits;
  for (u32 t0 = 0; t0 < im.n_tiles; t0 += 32) {
    const u32 t = t0 + lane;
    const u64 v = t < im.n_tiles ? P.tile_bits[im.tile_off + t] : 0;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u64 u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc += u; }
    if (t < im.n_tiles) P.tile_bitoff[im.tile_off + t] = run + inc - v;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) { P.st[ii].total_bits = run; P.st[ii].bytes = (u32)((run + 7) >> 3); }
}

// Pack the codes of a tile in shared memory, then move the words out (first / last word shared with the neighbours: OR).
constexpr int LL_EMIT_WORDS = (LL_TILE * LL_MAX_PIXEL_BITS + 31) / 32 + 2;
__device__ __forceinline__ void ll_smem_put(u32* buf, u32 pos, u64 bits, u32 n) {  // n <= 60
  if (n == 0) return;
  const u32 k = pos >> 5, sh = pos & 31;
  atomicOr(&buf[k], (u32)(bits << sh));
  if (sh + n > 32) {
    const u64 rest = bits >> (32 - sh);
    atomicOr(&buf[k + 1], (u32)rest);
    if (sh + n > 64) atomicOr(&buf[k + 2], (u32)(rest >> 32));
  }
}
__global__ void __launch_bounds__(LL_THREADS) k_ll_emit(LlParams P) {
  __shared__ u32 s_tab[4 * 280];
  __shared__ u32 s_buf[LL_EMIT_WORDS];
  __shared__ u32 s_red[9];
  const u32 tile = blockIdx.x;
  const u32 ii = ll_find_image(P.img, P.n_img, tile);
  const LlImage im = P.img[ii];
  for (u32 k = threadIdx.x; k < 4 * 280; k += LL_THREADS) s_tab[k] = P.codes[(size_t)ii * 4 * 280 + k];
  __syncthreads();
  auto table = [&](u32 ch, u32 sym) { return s_tab[ch * 280 + sym]; };
  const u32 p0 = (tile - im.tile_off) * LL_TILE + threadIdx.x * LL_PPT;
  const bool color = ll_is_color(im.color), alpha = ll_is_alpha(im.color);
  u32 px[LL_PPT], d[LL_PPT];
  ll_load4(P, im, p0, px, d);
  // every pixel's bit string once: literal (green, red, blue, alpha: api.rs:1115-1160) and run token
  u64 lit[LL_PPT];
  u32 lit_len[LL_PPT], run[LL_PPT], run_len[LL_PPT];
  u32 bits = 0;
#pragma unroll
  for (int j = 0; j < LL_PPT; j++) {
    lit[j] = 0; lit_len[j] = 0; run[j] = 0; run_len[j] = 0;
    if (d[j] & 1u) {
      u32 e = table(1, (px[j] >> 8) & 255u);
      u64 code = e & 0xFFFFu;
      u32 len = e >> 16;
      if (color) {
        e = table(0, px[j] & 255u); code |= (u64)(e & 0xFFFFu) << len; len += e >> 16;
        e = table(2, (px[j] >> 16) & 255u); code |= (u64)(e & 0xFFFFu) << len; len += e >> 16;
      }
      if (alpha) { e = table(3, px[j] >> 24); code |= (u64)(e & 0xFFFFu) << len; len += e >> 16; }
      lit[j] = code; lit_len[j] = len;
    }
    if (d[j] >> 1) {
      u32 sym, eb, ev;
      ll_run_symbol(d[j] >> 1, sym, eb, ev);
      const u32 e = table(1, sym);
      run[j] = (e & 0xFFFFu) | (ev << (e >> 16));
      run_len[j] = (e >> 16) + eb;
    }
    bits += lit_len[j] + run_len[j];
  }
  u32 tot;
  const u64 tile_bit = P.tile_bitoff[tile];
  const u32 skew = (u32)(tile_bit & 31);
  u32 pos = skew + ll_block_sum_excl(bits, s_red, tot);
  const u32 nw = (skew + tot + 31) >> 5;
  for (u32 k = threadIdx.x; k < nw + 2 && k < (u32)LL_EMIT_WORDS; k += LL_THREADS) s_buf[k] = 0;  // only the words this tile fills
  __syncthreads();
#pragma unroll
  for (int j = 0; j < LL_PPT; j++) {
    ll_smem_put(s_buf, pos, lit[j], lit_len[j]);
    pos += lit_len[j];
    ll_smem_put(s_buf, pos, (u64)run[j], run_len[j]);
    pos += run_len[j];
  }
  __syncthreads();
  u32* out = reinterpret_cast<u32*>(P.out + P.out_off[ii]);
  const u64 w0 = tile_bit >> 5;
  for (u32 k = threadIdx.x; k < nw; k += LL_THREADS) {
    const u32 v = s_buf[k];
    if (k == 0 || k == nw - 1) { if (v) atomicOr(&out[w0 + k], v); }
    else out[w0 + k] = v;
  }
  if (tile == im.tile_off) {  // the image's first tile also places the header bits and the container wrap
    const u32 hb = P.st[ii].hdr_bits, hw = (hb + 31) >> 5;
    const u32* hdr = P.hdr + (size_t)ii * LL_HDR_WORDS;
    for (u32 k = threadIdx.x; k < hw; k += LL_THREADS) { const u32 v = hdr[k]; if (v) atomicOr(&out[k], v); }
    if (P.container && threadIdx.x == 0) {  // RIFF, chunk_size(frame) + 4, WEBP, VP8L, len (api.rs:1224-1241, :1325-1329)
      const u32 len = P.st[ii].bytes;
      u32* h = out - 5;
      h[0] = 0x46464952u; h[1] = ((len + 1) & ~1u) + 8 + 4; h[2] = 0x50424557u; h[3] = 0x4C385056u; h[4] = len;
    }
  }
}
#endif  // __CUDACC__

}  // namespace zw
#endif
