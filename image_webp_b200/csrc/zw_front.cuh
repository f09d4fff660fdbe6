// zw_front.cuh -- front-end kernels: RGB->YUV420, macroblock analysis, segment assignment.
#ifndef ZW_FRONT_CUH
#define ZW_FRONT_CUH
#include "zw_types.cuh"

namespace zw {

// ---------------------------------------------------------------------------------------------
// (1) RGB(A) -> padded YUV 4:2:0 planes.   Reference: convert_image_yuv, src/decoder/yuv.rs:656-804
//     (rgb_to_y :859, rgb_to_u/v_avg :866-886, raw :889-899), 16-bit fixed point, 2x2 box
//     average with edge duplication, replicate padding to 16*mbw x 16*mbh.
//
// One warp converts a strip of 2 source rows x YUV_TILE_W pixels per step (8 warps per CTA, 2 steps).
// The packed RGB bytes of both rows are staged in shared memory with 128-bit loads (3 B/px is never
// 16-byte aligned per pixel, so threads pick their pixels out of the staged span); each thread then
// produces an 8x2 luma patch (two 64-bit stores) and 4 U + 4 V samples (two 32-bit stores).
// HBM-bound: 3 B/px read, 1.5 B/px written.
// ---------------------------------------------------------------------------------------------
constexpr int YUV_TILE_W = 256;                 // luma pixels per CTA strip
constexpr int YUV_THREADS = YUV_TILE_W / 8;     // 32 threads (one warp) per row pair, 8 px each
constexpr int YUV_ROWPAIRS = 8;                 // row pairs per CTA step (blockDim.y)
#ifndef ZW_YUV_STEPS
#define ZW_YUV_STEPS 8
#endif
constexpr int YUV_STEPS = ZW_YUV_STEPS;                    // steps per CTA: 128 source rows x 256 px; step s+1 is in flight (cp.async) while step s is converted
constexpr int YUV_ROW_SLOTS = (YUV_TILE_W * 4 + 32) / 16;  // uint4 slots per staged row (RGBA worst case + skew)

// dp2a: d = c + a.lo16 * b.byte(0|2) + a.hi16 * b.byte(1|3); unsigned or signed 16-bit coefficients x unsigned bytes
__device__ __forceinline__ u32 dp2a_lo_uu(u32 a, u32 b, u32 c) { u32 d; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ u32 dp2a_hi_uu(u32 a, u32 b, u32 c) { u32 d; asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ i32 dp2a_lo_su(i32 a, u32 b, i32 c) { i32 d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ i32 dp2a_hi_su(i32 a, u32 b, i32 c) { i32 d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

#ifndef ZW_YUV_MIN_BLOCKS
#define ZW_YUV_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(YUV_THREADS* YUV_ROWPAIRS, ZW_YUV_MIN_BLOCKS) k_yuv(ChunkParams P) {
  const ImageDesc d = P.img[blockIdx.z];
  const int pw = d.mbw * 16, ph = d.mbh * 16;
  const int x0 = blockIdx.x * YUV_TILE_W;
  if (x0 >= pw || (int)blockIdx.y * YUV_ROWPAIRS * YUV_STEPS * 2 >= ph) return;
  const int w = d.width, h = d.height, bpp = d.bpp;
  const int cw = (w + 1) >> 1, chh = (h + 1) >> 1;
  // span of source columns this strip needs (always non-empty: x0 < w because padding < 16)
  const int xs = x0, xe = min(x0 + YUV_TILE_W, w);
  const int span_bytes = (xe - xs) * bpp;
  extern __shared__ uint4 smem4[];
  u8* sm = reinterpret_cast<u8*>(smem4);
  u8* yp = P.planes + d.y_off;
  u8* up = yp + (size_t)pw * ph;
  u8* vp = up + (size_t)(pw >> 1) * (ph >> 1);
  const int tx = x0 + threadIdx.x * 8;
  // every warp (threadIdx.y) owns its staged rows (two buffers of two rows): no CTA-wide barrier, just __syncwarp.
  // The rows of step s+1 are fetched with 16-byte cp.async (LDGSTS: no registers held while the bytes are in flight)
  // before step s is converted, so each warp keeps two steps' worth of loads outstanding.
  uint4* stage0 = smem4 + (threadIdx.y * 4) * YUV_ROW_SLOTS;
  auto rows_of = [&](int step, int& rp, int& rowA, int& rowB) {
    rp = (blockIdx.y * YUV_STEPS + step) * YUV_ROWPAIRS + threadIdx.y;  // row pair in the padded plane
    const int ccy = min(rp, chh - 1);
    rowA = 2 * ccy; rowB = min(2 * ccy + 1, h - 1);
  };
  // 32-bit offsets inside the image (w * h * bpp < 2^32; every image starts 16-byte aligned in the arena)
  const u8* img_rgb = P.rgb + d.rgb_off;
  const u32 row_bytes = (u32)w * (u32)bpp, xs_bytes = (u32)xs * (u32)bpp;
  const u32 stage_sm = (u32)__cvta_generic_to_shared(stage0);
  auto issue = [&](int step) {
    int rp, rowA, rowB;
    rows_of(step, rp, rowA, rowB);
    if (step < YUV_STEPS && rp * 2 < ph) {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const u32 byte0 = (u32)(r == 0 ? rowA : rowB) * row_bytes + xs_bytes;
        const u32 al = byte0 & ~15u;
        const int nvec = ((int)(byte0 - al) + span_bytes + 15) >> 4;
        const u8* src = img_rgb + al;
        const u32 dst = stage_sm + (u32)(((step & 1) * 2 + r) * YUV_ROW_SLOTS * 16);
        for (int i = threadIdx.x; i < nvec; i += YUV_THREADS)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (u32)i), "l"(src + 16u * (u32)i) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
#pragma unroll 1
  for (int step = 0; step < YUV_STEPS; step++) {
    int rp, rowA, rowB;
    rows_of(step, rp, rowA, rowB);
    if (rp * 2 >= ph) break;
    issue(step + 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    int skew[2];
#pragma unroll
    for (int r = 0; r < 2; r++) skew[r] = (int)(((u32)(r == 0 ? rowA : rowB) * row_bytes + xs_bytes) & 15u);
    __syncwarp();
    if (tx < pw) {
      const u8* sA = sm + (size_t)(threadIdx.y * 4 + (step & 1) * 2 + 0) * YUV_ROW_SLOTS * 16 + skew[0];
      const u8* sB = sm + (size_t)(threadIdx.y * 4 + (step & 1) * 2 + 1) * YUV_ROW_SLOTS * 16 + skew[1];
      // luma rows 2rp and 2rp+1: source rows min(2rp,h-1) and min(2rp+1,h-1)
      const bool lumaA_is_A = (rp <= chh - 1);  // else row h-1 == rowB
      u32 y0w[2] = {0, 0}, y1w[2] = {0, 0}, uw = 0, vw = 0;
      if (bpp == 3 && tx + 8 <= w && 2 * rp + 1 < h) {
        // interior fast path: the thread's 8 pixels of both rows are 24 contiguous bytes each; fetch
        // them as 32-bit words (7 per row, funnel-shifted to byte alignment) instead of 96 byte loads
        u32 a[2][6];
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const u8* base = r == 0 ? sA : sB;
          const u32 o = (u32)(base - sm) + 24u * threadIdx.x;
          const u32* wp = reinterpret_cast<const u32*>(sm) + (o >> 2);
          const u32 sh = (o & 3u) * 8u;
          u32 wv[7];
#pragma unroll
          for (int k = 0; k < 7; k++) wv[k] = wp[k];
#pragma unroll
          for (int k = 0; k < 6; k++) a[r][k] = __funnelshift_r(wv[k], wv[k + 1], sh);
        }
        // Packed arithmetic: one PRMT gathers a pixel's [R, G, B, x] bytes, two dp2a (16-bit coefficients x bytes) give
        // its luma; the result is below 2^24 with Y exactly in byte 2, so four pixels are packed with three more PRMTs.
        // Chroma accumulates the same dp2a pair over the four pixels of a 2x2 block (the reference's sums are linear).
        // The unpacked form of this path (byte extraction + IMAD) made the kernel issue-bound at 69 % of the HBM peak.
        u32 pa[8], pb[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int w0 = (3 * k) >> 2, o = (3 * k) & 3, w1 = w0 + 1 < 6 ? w0 + 1 : w0;
          const u32 sel = (u32)o | ((u32)(o + 1) << 4) | ((u32)(o + 2) << 8) | ((u32)(o + 2) << 12);  // byte 3: don't care
          pa[k] = __byte_perm(a[0][w0], a[0][w1], sel);
          pb[k] = __byte_perm(a[1][w0], a[1][w1], sel);
        }
        const u32 cyRG = 16839u | (33059u << 16), cyB = 6420u, cyK = 32768u + (16u << 16);
        u32 ya[8], yb[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          ya[k] = dp2a_hi_uu(cyB, pa[k], dp2a_lo_uu(cyRG, pa[k], cyK));
          yb[k] = dp2a_hi_uu(cyB, pb[k], dp2a_lo_uu(cyRG, pb[k], cyK));
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
          y0w[h2] = __byte_perm(__byte_perm(ya[4 * h2], ya[4 * h2 + 1], 0x0062), __byte_perm(ya[4 * h2 + 2], ya[4 * h2 + 3], 0x0062), 0x5410);
          y1w[h2] = __byte_perm(__byte_perm(yb[4 * h2], yb[4 * h2 + 1], 0x0062), __byte_perm(yb[4 * h2 + 2], yb[4 * h2 + 3], 0x0062), 0x5410);
        }
        const i32 cuRG = (i32)((u32)(u16)(i16)-9719 | ((u32)(u16)(i16)-19081 << 16)), cuB = 28800;
        const i32 cvRG = (i32)((u32)(u16)(i16)28800 | ((u32)(u16)(i16)-24116 << 16)), cvB = (i32)(u32)(u16)(i16)-4684;
        const i32 cK = 4 * (128 << 16) + (32768 << 2);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          i32 u = cK, v = cK;
          const u32 q[4] = {pa[2 * k], pa[2 * k + 1], pb[2 * k], pb[2 * k + 1]};
#pragma unroll
          for (int t = 0; t < 4; t++) {
            u = dp2a_hi_su(cuB, q[t], dp2a_lo_su(cuRG, q[t], u));
            v = dp2a_hi_su(cvB, q[t], dp2a_lo_su(cvRG, q[t], v));
          }
          uw |= (u32)(u >> 18) << (8 * k);
          vw |= (u32)(v >> 18) << (8 * k);
        }
      } else if (bpp <= 2) {
        // L8 / La8 (convert_image_y, src/decoder/yuv.rs:806-845): Y = the grey sample, U = V = 127
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int sx = min(tx + k, w - 1) - xs;
          y0w[k >> 2] |= (u32)((lumaA_is_A ? sA : sB)[sx * bpp]) << (8 * (k & 3));
          y1w[k >> 2] |= (u32)(sB[sx * bpp]) << (8 * (k & 3));
        }
        uw = 0x7f7f7f7fu;
        vw = 0x7f7f7f7fu;
      } else {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int sx = min(tx + k, w - 1) - xs;
        const u8* pa = (lumaA_is_A ? sA : sB) + sx * bpp;
        const u8* pb = sB + sx * bpp;
        const int ya = (16839 * pa[0] + 33059 * pa[1] + 6420 * pa[2] + 32768 + (16 << 16)) >> 16;
        const int yb = (16839 * pb[0] + 33059 * pb[1] + 6420 * pb[2] + 32768 + (16 << 16)) >> 16;
        y0w[k >> 2] |= (u32)ya << (8 * (k & 3));
        y1w[k >> 2] |= (u32)yb << (8 * (k & 3));
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int ccx = min((tx >> 1) + k, cw - 1);
        const int c0 = 2 * ccx - xs, c1 = min(2 * ccx + 1, w - 1) - xs;
        const u8 *p1 = sA + c0 * bpp, *p2 = sA + c1 * bpp, *p3 = sB + c0 * bpp, *p4 = sB + c1 * bpp;
        const int r = p1[0] + p2[0] + p3[0] + p4[0];
        const int g = p1[1] + p2[1] + p3[1] + p4[1];
        const int b = p1[2] + p2[2] + p3[2] + p4[2];
        const int u = (-9719 * r - 19081 * g + 28800 * b + 4 * (128 << 16) + (32768 << 2)) >> 18;
        const int v = (28800 * r - 24116 * g - 4684 * b + 4 * (128 << 16) + (32768 << 2)) >> 18;
        uw |= (u32)u << (8 * k);
        vw |= (u32)v << (8 * k);
      }
      }
      const u32 yo = (u32)(2 * rp) * (u32)pw + (u32)tx, co = (u32)rp * (u32)(pw >> 1) + (u32)(tx >> 1);
      *reinterpret_cast<uint2*>(yp + yo) = make_uint2(y0w[0], y0w[1]);
      *reinterpret_cast<uint2*>(yp + yo + (u32)pw) = make_uint2(y1w[0], y1w[1]);
      *reinterpret_cast<u32*>(up + co) = uw;
      *reinterpret_cast<u32*>(vp + co) = vw;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// (2) Analysis pass: per-macroblock "alpha" from DCT histograms of DC and TM predictions built
//     from SOURCE neighbours (no reconstruction dependency -> every macroblock is independent).
//     Reference: analyze_image src/encoder/analysis.rs:964, AnalysisIterator::import :622,
//     make_luma16_preds :365, make_chroma8_preds :460, forward_dct_4x4 :172,
//     collect_histogram_with_offset :922, DctHistogram :139-166, analyze_macroblock :951.
//     One warp per macroblock: lanes 0..23 own one 4x4 block each (16 Y, 4 U, 4 V) and transform
//     it against both predictions; 32-bin histograms live in shared memory.
// ---------------------------------------------------------------------------------------------
#ifndef ZW_AN_WARPS
#define ZW_AN_WARPS 2  // measured 8 -> 4 -> 2 warps per CTA: 2.43 -> 2.33 -> 2.31 ms (small CTAs free their slots early)
#endif
constexpr int AN_WARPS = ZW_AN_WARPS;

__device__ __forceinline__ void an_fdct_hist(const i32* res, u32* hist) {
  i32 c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = res[i];
  fdct4x4(c);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    // the reference narrows to i16 before |.| >> 3 (analysis.rs:204-207, :236); values fit i16
    int v = imin(iabs(c[i]) >> 3, 31);
    atomicAdd(&hist[v], 1u);
  }
}

__global__ void __launch_bounds__(AN_WARPS * 32) k_analysis(ChunkParams P) {
  __shared__ u32 s_hist[AN_WARPS][4][32];  // [warp][y-dc, y-tm, uv-dc, uv-tm][bin]
  __shared__ __align__(16) u8 s_tile[AN_WARPS][17 * 32 + 2 * 9 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 gmb = blockIdx.x * AN_WARPS + warp;
  const int img = blockIdx.y;
  const ImageDesc d = P.img[img];
  const u32 nmb = d.mbw * d.mbh;
  const bool valid = d.use_segments && gmb < nmb;
  for (int i = lane; i < 128; i += 32) (&s_hist[warp][0][0])[i] = 0;
  __syncwarp();
  if (valid) {
    const int mbx = gmb % d.mbw, mby = gmb / d.mbw;
    const int pw = d.mbw * 16, cwid = d.mbw * 8;
    const u8* yp = P.planes + d.y_off;
    const u8* up = yp + (size_t)pw * d.mbh * 16;
    const u8* vp = up + (size_t)cwid * d.mbh * 8;
    const bool has_left = mbx > 0, has_top = mby > 0;
    // stage the macroblock with its top row / left column / corner in shared memory: one 16-byte (luma) or
    // 8-byte (chroma) row per lane instead of ~50 strided byte loads per lane from global memory
    u8* ty = s_tile[warp];            // [17][32]: row 0 = the row above, column 15 = the column to the left, pixels at 16..31
    u8* tu = ty + 17 * 32;            // [9][16]:  column 7 = left, pixels at 8..15
    u8* tv = tu + 9 * 16;
    if (lane < 17) {
      const int r = lane - 1;         // source row relative to the macroblock
      if (r >= 0 || has_top) {
        const u8* src = yp + (size_t)(mby * 16 + r) * pw + mbx * 16;
        *reinterpret_cast<uint4*>(&ty[lane * 32 + 16]) = __ldg(reinterpret_cast<const uint4*>(src));
        if (has_left) ty[lane * 32 + 15] = __ldg(src - 1);
      }
    } else if (lane < 26) {
      const int r = lane - 18;
      if (r >= 0 || has_top) {
        const u8* src = up + (size_t)(mby * 8 + r) * cwid + mbx * 8;
        *reinterpret_cast<uint2*>(&tu[(lane - 17) * 16 + 8]) = __ldg(reinterpret_cast<const uint2*>(src));
        if (has_left) tu[(lane - 17) * 16 + 7] = __ldg(src - 1);
      }
    }
    if (lane < 9) {  // V rows (lanes 0..8 a second time)
      const int r = lane - 1;
      if (r >= 0 || has_top) {
        const u8* src = vp + (size_t)(mby * 8 + r) * cwid + mbx * 8;
        *reinterpret_cast<uint2*>(&tv[lane * 16 + 8]) = __ldg(reinterpret_cast<const uint2*>(src));
        if (has_left) tv[lane * 16 + 7] = __ldg(src - 1);
      }
    }
    __syncwarp();
    if (lane < 24) {
      const u8* o;   // macroblock origin inside its tile
      int stride, bx, by, size;
      if (lane < 16) { o = ty + 32 + 16; stride = 32; bx = lane & 3; by = lane >> 2; size = 16; }
      else { o = (lane < 20 ? tu : tv) + 16 + 8; stride = 16; bx = lane & 1; by = (lane >> 1) & 1; size = 8; }
      // DC value over the whole MB border (analysis.rs:259-291 / :378-419)
      int dcv;
      {
        int st = 0, sl = 0;
        if (has_top) for (int i = 0; i < size; i++) st += o[i - stride];
        if (has_left) for (int i = 0; i < size; i++) sl += o[i * stride - 1];
        const int shift = size == 16 ? 5 : 4;
        if (has_top && has_left) dcv = (st + sl + size) >> shift;
        else if (has_top) dcv = (2 * st + size) >> shift;
        else if (has_left) dcv = (2 * sl + size) >> shift;
        else dcv = 0x80;
      }
      // corner (analysis.rs:594, :676-684): 127 on row 0 else the source pixel; x==0 never uses it
      const int tl = has_top ? (has_left ? o[-stride - 1] : 129) : 127;
      i32 rdc[16], rtm[16];
#pragma unroll
      for (int y = 0; y < 4; y++)
#pragma unroll
        for (int x = 0; x < 4; x++) {
          const int xx = bx * 4 + x, yy = by * 4 + y;
          const int s = o[yy * stride + xx];
          int tm;
          if (has_left && has_top) tm = clip255((int)o[yy * stride - 1] + (int)o[xx - stride] - tl);
          else if (has_left) tm = o[yy * stride - 1];   // horizontal_pred
          else if (has_top) tm = o[xx - stride];         // vertical_pred
          else tm = 129;
          rdc[y * 4 + x] = s - dcv;
          rtm[y * 4 + x] = s - tm;
        }
      const int hsel = lane < 16 ? 0 : 2;
      an_fdct_hist(rdc, s_hist[warp][hsel]);
      an_fdct_hist(rtm, s_hist[warp][hsel + 1]);
    }
    __syncwarp();
    // DctHistogram::from_distribution + get_alpha for the 4 histograms: lane = bin
    int alpha[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 cnt = s_hist[warp][k][lane];
      u32 mx = cnt;
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
      const u32 nzmask = __ballot_sync(0xffffffffu, cnt > 0);
      const int last_nz = nzmask ? 31 - __clz(nzmask) : 1;
      alpha[k] = mx > 1 ? (int)(510u * (u32)last_nz / mx) : 0;
    }
    if (lane == 0) {
      const int best = max(max(alpha[0], alpha[1]), -1);
      const int best_uv = max(max(alpha[2], alpha[3]), -1);
      int a = (3 * best + best_uv + 2) >> 2;
      a = clip255(255 - a);
      P.alpha[d.mb_off + gmb] = (u8)a;
      atomicAdd(&P.alpha_hist[(size_t)img * 256 + a], 1u);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (3) Segments: 1-D k-means over the 256-bin alpha histogram, per-segment quantiser index,
//     segment map, segment-tree probabilities.  One CTA per image; thread 0 runs the (tiny,
//     strictly sequential) integer k-means, all threads map macroblocks.
//     Reference: assign_segments_kmeans src/encoder/analysis.rs:1029-1130;
//     analyze_and_assign_segments src/encoder/vp8.rs:2278-2388; compute_segment_quant
//     analysis.rs:1145-1174 (f64 pow -> host-built LUT, bit-exact, see zw_capi.cu).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_segments(ChunkParams P) {
  const int img = blockIdx.x;
  const ImageDesc d = P.img[img];
  ImageState& S = P.st[img];
  __shared__ u8 s_map[256];
  __shared__ u32 s_cnt[4];
  if (!d.use_segments) {
    if (threadIdx.x == 0) {
      S.seg_enabled = 0; S.update_map = 0;
      for (int i = 0; i < 4; i++) { S.seg_qidx[i] = (u8)P.base_qidx; S.seg_delta[i] = 0; }
      S.tree_probs[0] = S.tree_probs[1] = S.tree_probs[2] = 255;
    }
    return;
  }
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    const u32* alphas = P.alpha_hist + (size_t)img * 256;
    int centers[4] = {0, 0, 0, 0};
    for (int i = 0; i < 256; i++) s_map[i] = 0;
    int min_a = 0, max_a = 255;
    for (int n = 0; n < 256; n++) if (alphas[n] > 0) { min_a = n; break; }
    for (int n = 255; n >= min_a; n--) if (alphas[n] > 0) { max_a = n; break; }
    const int range_a = max_a >= min_a ? max_a - min_a : 0;
    for (int k = 0; k < 4; k++) centers[k] = (min_a + ((1 + 2 * k) * range_a) / 8) & 255;
    u32 accum[4], dist_accum[4];
    int weighted_average = 0;
    u32 total_weight = 0;
    for (int iter = 0; iter < 6; iter++) {
      for (int i = 0; i < 4; i++) { accum[i] = 0; dist_accum[i] = 0; }
      int cur = 0;
      for (int a = min_a; a <= max_a; a++) {
        if (alphas[a] > 0) {
          while (cur + 1 < 4) {
            const int dc = iabs(a - centers[cur]), dn = iabs(a - centers[cur + 1]);
            if (dn < dc) cur++; else break;
          }
          s_map[a] = (u8)cur;
          dist_accum[cur] += (u32)a * alphas[a];
          accum[cur] += alphas[a];
        }
      }
      int displaced = 0;
      weighted_average = 0;
      total_weight = 0;
      for (int n = 0; n < 4; n++) {
        if (accum[n] > 0) {
          const int nc = (int)((dist_accum[n] + accum[n] / 2) / accum[n]) & 255;
          displaced += iabs(centers[n] - nc);
          centers[n] = nc;
          weighted_average += nc * (int)accum[n];
          total_weight += accum[n];
        }
      }
      if (displaced < 5) break;
    }
    if (total_weight > 0) weighted_average = (weighted_average + (int)total_weight / 2) / (int)total_weight;
    else weighted_average = 128;
    int mn = 255, mx = 0;
    for (int k = 0; k < 4; k++) { mn = min(mn, centers[k]); mx = max(mx, centers[k]); }
    const int range = mx == mn ? 1 : mx - mn;
    for (int k = 0; k < 4; k++) {
      int ta = 255 * (centers[k] - weighted_average) / range;  // truncating division (vp8.rs:2326)
      ta = imin(imax(ta, -127), 127);
      const int q = P.segquant_lut[P.base_qidx * 255 + (ta + 127)];
      S.seg_qidx[k] = (u8)q;
      S.seg_delta[k] = (i8)(q - (int)P.base_qidx);
      S.centers[k] = (u8)centers[k];
    }
    S.mid_alpha = weighted_average;
    for (int i = 0; i < 256; i++) P.map256[(size_t)img * 256 + i] = s_map[i];
  }
  __syncthreads();
  const u32 nmb = d.mbw * d.mbh;
  u32 local[4] = {0, 0, 0, 0};
  for (u32 i = threadIdx.x; i < nmb; i += blockDim.x) {
    const u8 s = s_map[P.alpha[d.mb_off + i]];
    P.segmap[d.mb_off + i] = s;
    local[s]++;
  }
  for (int k = 0; k < 4; k++) if (local[k]) atomicAdd(&s_cnt[k], local[k]);
  __syncthreads();
  if (threadIdx.x == 0) {
    auto proba = [](u32 a, u32 b) -> u8 {
      const u32 t = a + b;
      return t == 0 ? 255 : (u8)((255 * a + t / 2) / t);
    };
    S.tree_probs[0] = proba(s_cnt[0] + s_cnt[1], s_cnt[2] + s_cnt[3]);
    S.tree_probs[1] = proba(s_cnt[0], s_cnt[1]);
    S.tree_probs[2] = proba(s_cnt[2], s_cnt[3]);
    S.update_map = (S.tree_probs[0] != 255 || S.tree_probs[1] != 255 || S.tree_probs[2] != 255) ? 1 : 0;
    S.seg_enabled = 1;
    for (int k = 0; k < 4; k++) S.seg_count[k] = s_cnt[k];
  }
}

}  // namespace zw
#endif
