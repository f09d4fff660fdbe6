// zw_searchq.cuh -- the luma wavefront kernels with FOUR lanes per macroblock row (zw_quad.cuh): a warp walks eight
// rows of (normally) eight different images at once, every 4x4 block / I4 candidate / trellis lane-private.
// Scheduling, dependency flags and the record layout are those of k_search (zw_search.cuh); only the mapping of
// the work inside a macroblock differs.  Used when the chunk has enough rows to fill the GPU with quads (batches);
// single images keep the warp-per-row kernels.
#ifndef ZW_SEARCHQ_CUH
#define ZW_SEARCHQ_CUH
#include "zw_quad.cuh"
#include "zw_search.cuh"

namespace zw {

// One CTA per SM.  Pass 1 (all level costs zero, light on registers): 12 warps = 96 macroblock rows in flight, ~222 KB of
// shared memory, 170 registers (measured 17.5 -> 16.6 ms against 8 warps).  Pass 2 (real cost tables; methods 2-3 only):
// 8 warps = 64 rows, up to 255 registers -- 12 warps were measured slower there.
#ifndef ZW_SQ_WARPS1
#define ZW_SQ_WARPS1 12
#endif
#ifndef ZW_SQ_WARPS2
#define ZW_SQ_WARPS2 8
#endif
#ifndef ZW_SQ_MIN_BLOCKS
#define ZW_SQ_MIN_BLOCKS 1
#endif
__host__ __device__ constexpr int sq_warps(int pass) { return pass == 1 ? ZW_SQ_WARPS1 : ZW_SQ_WARPS2; }  // warps per CTA (8 quads each)
__host__ __device__ constexpr int sq_quads(int pass) { return sq_warps(pass) * 8; }

struct SearchQShared {
  u8 pred_idx[10][16];
  u16 dtaps[32];
  QuadScratch q[1];  // sq_quads(pass) entries (dynamic shared memory)
};

struct QuadExecDev {
  int q;
  unsigned mask;
  template <class F>
  __device__ __forceinline__ void run(F&& f) {
    __syncwarp(mask);  // everybody has finished reading what the previous step left in shared memory
    f(q);
    __syncwarp(mask);
  }
};

// The CTA's warps step through the three phases of a macroblock (I16 search, I4 search, final transform) in LOCK STEP, one
// CTA barrier per phase: the kernel is ~100 KB of code against a 32 KB instruction cache per SM, and only the loop body
// of one phase fits.  A quad whose dependency is not ready (it never blocks inside a round) or that has no row left sits
// the round out.
template <int PASS>
__global__ void __launch_bounds__(sq_warps(PASS) * 32, ZW_SQ_MIN_BLOCKS) k_searchq(ChunkParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SearchQShared& SH = *reinterpret_cast<SearchQShared*>(smem_raw);
  __shared__ int s_active;  // quads of this CTA that still have (or may get) a row
  for (int i = threadIdx.x; i < 160; i += blockDim.x) (&SH.pred_idx[0][0])[i] = (&d_pred_idx[0][0])[i];
  if (threadIdx.x < 32) SH.dtaps[threadIdx.x] = d_dtaps[threadIdx.x];
  if (threadIdx.x == 0) s_active = sq_quads(PASS);
  __syncthreads();
  for (int i = threadIdx.x; i < 16; i += blockDim.x) SH.pred_idx[1][i] = (u8)(32 + i);  // TM pixels live in dtab[32 + n]
  __syncthreads();
  const int lane = threadIdx.x & 31, q = lane & 3, qbase = lane & 28;
  const unsigned qmask = 0xFu << qbase;
  QuadScratch& S = SH.q[threadIdx.x >> 2];
  QuadExecDev X;
  X.q = q; X.mask = qmask;
  QuadConst K;
  K.pidx = SH.pred_idx; K.dtaps = SH.dtaps;
  int* progress = P.progress + (PASS - 1) * P.n_rows;
  MbRecord* recs = PASS == 1 ? P.rec1 : P.rec2;
  const bool trellis = (PASS == 2) && P.do_trellis;

  // state of the row this quad owns
  bool have_row = false, done = false, seg_on = false;
  u32 img = 0, row_mb0 = 0, up_mb0 = 0, row_off = 0, left_nz = 0;
  int mbw = 0, mby = 0, mbx = 0, seen = 0;
  const u8* yp = nullptr;
  QuadMbIn in;
  in.cc.probs = nullptr; in.cc.level_cost = nullptr; in.SP = &P.segtab[P.base_qidx];
  in.i4_modes = (int)P.i4_modes; in.i4_always = P.i4_always != 0; in.trellis = trellis;
  in.mbx = 0; in.mby = 0; in.in_top_nz = 0; in.in_left_nz = 0;
  QuadMbState st;
  st.dc16 = 0; st.best16_mode = 0; st.i16_score = 0; st.use_i4 = false;

  for (;;) {
    if (!have_row && !done) {
      u32 t = 0;
      if (q == 0) t = atomicAdd(&P.ticket[PASS - 1], 1u);
      t = __shfl_sync(qmask, t, qbase);
      if (t >= P.n_rows) {
        done = true;
        if (q == 0) atomicSub(&s_active, 1);
      } else {
        const RowRef rr = P.rows[t];
        const ImageDesc d = P.img[rr.img];
        img = rr.img; mbw = d.mbw; mby = rr.mby; row_off = d.row_off;
        yp = P.planes + d.y_off;
        row_mb0 = d.mb_off + mby * mbw;
        up_mb0 = row_mb0 - mbw;  // only dereferenced when mby > 0
        in.cc.probs = PASS == 1 ? ZW_TAB(kCoeffProbs) : P.probs + (size_t)img * 1056;
        in.cc.level_cost = PASS == 1 ? nullptr : P.lcost + (size_t)img * 6528;
        in.mby = mby;
        seg_on = P.st[img].seg_enabled != 0;
        left_nz = 0; mbx = 0; seen = 0;
        have_row = true;
      }
    }
    // ---- dependency: the top / top-right neighbours (row above finished mbx + 1); look once, never block ----
    bool work = have_row;
    if (work && mby > 0) {
      const int need = min(mbx + 2, mbw);
      int ok = 1;
      if (q == 0 && seen < need) {
        seen = ld_flag(&progress[row_off + mby - 1]);
        ok = seen >= need;
        if (ok) fence_acquire();
      }
      work = __shfl_sync(qmask, ok, qbase) != 0;
    }
    __syncthreads();           // phase barrier 1 (s_active only changes at ticket time, i.e. before it)
    if (s_active == 0) break;  // every thread reads the same value: the next change comes after barrier 3
    const u32 gmb = row_mb0 + mbx;
    int seg = 0;
    if (work) {
      const int pw = mbw * 16;
      seg = seg_on ? P.segmap[gmb] : 0;
      in.SP = &P.segtab[seg_on ? P.st[img].seg_qidx[seg] : P.base_qidx];
      in.mbx = mbx;
      in.in_top_nz = (PASS == 2 && mby > 0) ? (u32)__ldcg(&P.nz_after[up_mb0 + mbx]) : 0u;
      in.in_left_nz = left_nz;
      {  // stage the macroblock: source rows 4q .. 4q+3, borders, zeroed levels
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(yp + (size_t)(mby * 16 + 4 * q + r) * pw + mbx * 16));
          *reinterpret_cast<uint4*>(&S.src_y[(4 * q + r) * 16]) = v;
        }
        if (mby == 0) {
#pragma unroll
          for (int k = 0; k < 8; k++) S.yws[8 * q + k] = 127;  // corner + 16 above + top-right: all 127 on the first row
        } else {
          const MbBottom* bt = &P.bottom[up_mb0 + mbx];
          const u32 w = __ldcg(reinterpret_cast<const u32*>(&bt->y[4 * q]));
#pragma unroll
          for (int k = 0; k < 4; k++) S.yws[1 + 4 * q + k] = (u8)(w >> (8 * k));
          if (q == 3) {  // top-right 4 pixels: next MB's bottom row, or the replicated last pixel
            const u32 tr = (mbx == mbw - 1) ? (w >> 24) * 0x01010101u : __ldcg(reinterpret_cast<const u32*>(&(bt + 1)->y[0]));
#pragma unroll
            for (int k = 0; k < 4; k++) S.yws[17 + k] = (u8)(tr >> (8 * k));
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) S.yws[(1 + 4 * q + k) * 32] = (mbx == 0) ? (u8)129 : S.left_y[1 + 4 * q + k];
        if (q == 0 && mby != 0) S.yws[0] = mbx == 0 ? (u8)129 : S.left_y[0];
        for (int k = q; k < 136; k += QG) reinterpret_cast<u32*>(S.lv)[k] = 0;
      }
      __syncwarp(qmask);
      if (q < 3) {  // top-right copies for sub-block rows 1..3 (prediction.rs:49-54)
        const int r = 4 * (1 + q);
#pragma unroll
        for (int k = 0; k < 4; k++) S.yws[r * 32 + 17 + k] = S.yws[17 + k];
      }
      __syncwarp(qmask);
      quad_i16<1>(X, S, K, in, st);
    }
    __syncthreads();  // phase barrier 2
    if (work) quad_i4<1>(X, S, K, in, st);
    __syncthreads();  // phase barrier 3
    if (!work) continue;
    const QuadLumaOut L = quad_final<1>(X, S, K, in, st);
    __syncwarp(qmask);
    bool skip = false;
    u32 out_top = 0, out_left = 0;
    if (PASS == 2) {
      const u32 uvnz = P.uvflags[gmb];  // chroma of this macroblock was coded by k_chroma2 (it does not depend on luma)
      skip = !(L.simple_nz || uvnz != 0);
      q_complexity_after(L.use_i4, skip, L.y2nz, L.ynz, uvnz, in.in_top_nz, left_nz, out_top, out_left);
    }
    {
      u32* g = reinterpret_cast<u32*>(&recs[gmb]);
      if (q == 0) {
        // header: ymode, uvmode, segment, skip | 16 sub-block modes | top_nz, left_nz | derr_left | derr_top
        const u32 ym = L.use_i4 ? 4u : (u32)L.mode16;
        if (PASS == 2) {
          const u32 uvm = (g[0] >> 8) & 255u;  // chroma-owned header fields: keep what k_chroma2 stored (also words 6, 7)
          g[0] = ym | (uvm << 8) | ((u32)seg << 16) | ((u32)skip << 24);
          g[5] = (in.in_top_nz & 0xffffu) | (left_nz << 16);
        } else {
          // pass 1: luma flags parked here until k_finish1 completes the record
          g[0] = ym | ((u32)seg << 16);
          g[5] = (L.ynz & 0xffffu) | (((L.y2nz ? 1u : 0u) | (L.simple_nz ? 2u : 0u)) << 16);
          g[6] = 0; g[7] = 0;
        }
      }
      g[1 + q] = L.use_i4 ? *reinterpret_cast<const u32*>(&S.bmodes[4 * q]) : 0u;
      if (PASS == 2 && skip) {
        for (int k = 8 + q; k < 208; k += QG) g[k] = 0u;  // a skipped MB codes nothing: zero all 200 level words
      } else {
        const u32* sl = reinterpret_cast<const u32*>(S.lv);
        for (int k = q; k < 136; k += QG) g[8 + k] = sl[k];  // chroma levels: k_chroma2 / the pass-1 chroma chain
      }
    }
    left_nz = out_left;
    // borders for the neighbours
    MbBottom* bo = &P.bottom[gmb];
    {
      u32 w = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) w |= (u32)S.yws[16 * 32 + 1 + 4 * q + k] << (8 * k);
      *reinterpret_cast<u32*>(&bo->y[4 * q]) = w;
    }
    u8 lcol[5];
#pragma unroll
    for (int k = 0; k < 4; k++) lcol[k] = S.yws[(1 + 4 * q + k) * 32 + 16];
    lcol[4] = S.yws[16];
    if (PASS == 2 && q == 0) P.nz_after[gmb] = (u16)out_top;
    __syncwarp(qmask);
#pragma unroll
    for (int k = 0; k < 4; k++) S.left_y[1 + 4 * q + k] = lcol[k];
    if (q == 0) { S.left_y[0] = lcol[4]; publish_flag(&progress[row_off + mby], mbx + 1); }
    mbx++;
    if (mbx == mbw) have_row = false;
  }
}

__host__ __device__ constexpr size_t searchq_smem_bytes(int pass) { return sizeof(SearchQShared) + (size_t)(sq_quads(pass) - 1) * sizeof(QuadScratch); }

}  // namespace zw
#endif
