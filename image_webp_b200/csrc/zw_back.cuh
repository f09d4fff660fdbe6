// zw_back.cuh -- token statistics, probability update, tokeniser, boolean coder, assembly.
//
// Reference (file:line under /root/reference):
//   record_residual_stats   src/encoder/vp8.rs:1027-1198      (k_rowstats)
//   ProbaStats::record      src/encoder/cost.rs:1200-1209      (ordered halving, k_probs fix-up)
//   compute_updated_probabilities vp8.rs:1202-1238, should_update cost.rs:1226-1254 (k_probs)
//   LevelCosts::calculate   cost.rs:1500-1546                  (k_probs)
//   skip probability        vp8.rs:1389-1395                   (k_probs)
//   encode_compressed_frame_header vp8.rs:332-372 (+ :393-496) (frame_header_tokens)
//   write_macroblock_header vp8.rs:498-560                     (mb_header_tokens)
//   encode_residual_data / encode_coefficients vp8.rs:650-958  (block_tokens)
//   ArithmeticEncoder       src/encoder/arithmetic.rs:7-196    (k_bc_*)
//   write_uncompressed_frame_header / write_partitions vp8.rs:315-330, :374-391 (k_assemble)
#ifndef ZW_BACK_CUH
#define ZW_BACK_CUH
#include "zw_boolcoder.cuh"
#include "zw_searchq.cuh"

namespace zw {

// ---- per-block contexts shared by the statistics and the tokeniser ----------------------------
// lane b (0..24) owns block b of the record: 0 Y2, 1..16 Y (raster), 17..20 U, 21..24 V.
struct BlockInfo {
  int type;    // token type / plane: 0 I16-AC, 1 I16-DC (Y2), 2 chroma, 3 I4
  int first;   // first coded coefficient
  int ctx;     // initial context (left + top has_coeffs)
  bool coded;  // block exists in this macroblock
};

__device__ __forceinline__ BlockInfo block_info(const MbRecord& r, int b, u32 nzmask) {
  BlockInfo I;
  const bool is_b = r.ymode == 4;
  const u32 tn = r.top_nz, ln = r.left_nz;
  I.coded = b < 25 && !(is_b && b == 0);
  if (b == 0) {
    I.type = 1; I.first = 0;
    I.ctx = (int)(ln & 1) + (int)(tn & 1);
  } else if (b <= 16) {
    const int x = (b - 1) & 3, y = (b - 1) >> 2;
    I.type = is_b ? 3 : 0; I.first = is_b ? 0 : 1;
    const int left = x > 0 ? (int)((nzmask >> (b - 1)) & 1) : (int)((ln >> (1 + y)) & 1);
    const int top = y > 0 ? (int)((nzmask >> (b - 4)) & 1) : (int)((tn >> (1 + x)) & 1);
    I.ctx = left + top;
  } else {
    const int c = b - 17, ch = c >> 2, x = c & 1, y = (c >> 1) & 1;
    I.type = 2; I.first = 0;
    const int sh = 5 + 2 * ch;
    const int left = x > 0 ? (int)((nzmask >> (b - 1)) & 1) : (int)((ln >> (sh + y)) & 1);
    const int top = y > 0 ? (int)((nzmask >> (b - 2)) & 1) : (int)((tn >> (sh + x)) & 1);
    I.ctx = left + top;
  }
  return I;
}

__device__ __forceinline__ bool block_nonzero(const i16* zz) {
  const uint4* p = reinterpret_cast<const uint4*>(zz);
  const uint4 a = p[0], b = p[1];
  return (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) != 0;
}

// ---------------------------------------------------------------------------------------------
// (4) Token statistics.  One warp per macroblock row accumulates (total, ones) for the 1056 slots
//     in shared memory with atomics (order-free), then stores them per row.  The order-dependent
//     halving (Q9) is reconstructed exactly in k_probs from these per-row counts.
// ---------------------------------------------------------------------------------------------
#ifndef ZW_STAT_WARPS
#define ZW_STAT_WARPS 2  // measured 4 -> 2 warps per CTA: 1.92 -> 1.85 ms
#endif
constexpr int STAT_WARPS = ZW_STAT_WARPS;

__global__ void __launch_bounds__(STAT_WARPS * 32) k_rowstats(ChunkParams P) {
  __shared__ u32 s_cnt[STAT_WARPS][1056 * 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 row = blockIdx.x * STAT_WARPS + warp;
  u32* cnt = s_cnt[warp];
  for (int i = lane; i < 2112; i += 32) cnt[i] = 0;
  __syncwarp();
  if (row < P.n_rows) {
    const RowRef rr = P.rows[row];
    const ImageDesc d = P.img[rr.img];
    const MbRecord* recs = P.rec1 + d.mb_off + (size_t)rr.mby * d.mbw;
    for (u32 mbx = 0; mbx < d.mbw; mbx++) {
      const MbRecord& r = recs[mbx];
      if (mbx + 1 < d.mbw) {  // the next record's header + this lane's levels on their way to L1 while this one is walked
        const MbRecord& nx = recs[mbx + 1];
        asm volatile("prefetch.global.L1 [%0];" ::"l"(lane < 25 ? (const void*)nx.levels[lane] : (const void*)&nx));
      }
      if (r.skip) continue;
      const bool nz = lane < 25 && block_nonzero(r.levels[lane < 25 ? lane : 0]);
      const u32 nzmask = __ballot_sync(FULL, nz);
      const BlockInfo I = block_info(r, lane, nzmask);
      if (I.coded) {
        token_events(r.levels[lane], I.type, I.first, I.ctx, [&](int slot, int bit) {
          atomicAdd(&cnt[slot * 2], 1u);
          if (bit) atomicAdd(&cnt[slot * 2 + 1], 1u);
        });
      }
      __syncwarp();
    }
    // canonical row index of this ticket: rows[] is a permutation, store by (image,row) position
    u32* out = P.rowstats + (size_t)(d.row_off + rr.mby) * 2112;
    __syncwarp();
    for (int i = lane; i < 2112; i += 32) out[i] = cnt[i];
  }
}

// Replay helper for the halving fix-up: (total, ones) of `slot` in one macroblock.
__device__ __forceinline__ void mb_slot_counts(const MbRecord& r, int slot, u32& tot, u32& ones, u32 limit) {
  // counts the first `limit` events of the slot in reference order
  tot = 0; ones = 0;
  if (r.skip) return;
  u32 nzmask = 0;
  for (int b = 0; b < 25; b++) nzmask |= (u32)block_nonzero(r.levels[b]) << b;
  for (int b = 0; b < 25; b++) {
    const BlockInfo I = block_info(r, b, nzmask);
    if (!I.coded) continue;
    token_events(r.levels[b], I.type, I.first, I.ctx, [&](int s, int bit) {
      if (s == slot && tot < limit) { tot++; ones += (u32)bit; }
    });
  }
}

// ---------------------------------------------------------------------------------------------
// (4b) Per image: exact ProbaStats (with ordered halving), probability update decision, level-cost
//      tables, skip probability.  One CTA per image.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_probs(ChunkParams P) {
  const int img = blockIdx.x;
  const ImageDesc d = P.img[img];
  ImageState& IS = P.st[img];
  __shared__ u32 s_stats[1056];
  __shared__ u8 s_newp[1056];
  __shared__ int s_sav[1056];
  __shared__ u32 s_total;  // i32 sum with wrap-around (release-mode Rust semantics), kept as u32
  __shared__ int s_nupd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) { s_total = 0; s_nupd = 0; }
  const u32* rs = P.rowstats + (size_t)d.row_off * 2112;
  // one warp per slot: lanes stride over rows
  for (int slot = warp; slot < 1056; slot += nwarp) {
    u32 tot = 0, ones = 0;
    for (u32 r = lane; r < d.mbh; r += 32) { tot += rs[(size_t)r * 2112 + slot * 2]; ones += rs[(size_t)r * 2112 + slot * 2 + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { tot += __shfl_xor_sync(FULL, tot, o); ones += __shfl_xor_sync(FULL, ones, o); }
    u32 packed;
    if (tot <= 65534u) {
      packed = (tot << 16) | ones;  // no halving ever triggered (cost.rs:1203 needs total >= 0xfffe before an event)
    } else {
      // Exact replay of the halving points.  Event index k (0-based) halves first when the running
      // total is 65534, i.e. at k = 65534 + 32767*j.  Ones between consecutive halving points are
      // counted from the per-row sums plus a replay of the one row (and macroblock) straddling it.
      u32 T = 0, O = 0;      // running packed state
      u32 done = 0;          // events consumed
      u32 row = 0, row_start = 0;  // first row not fully consumed, event index at its start
      u32 ones_before_row = 0;     // ones in rows [0,row)
      u32 consumed_ones = 0;       // ones in events [0,done)
      u32 boundary = 65534u;
      while (done < tot) {
        const u32 target = boundary < tot ? boundary : tot;  // consume events [done, target)
        // advance `row` so that target lies in (row_start, row_start+row_tot]
        u32 rt, ro;
        for (;;) {
          rt = rs[(size_t)row * 2112 + slot * 2];
          ro = rs[(size_t)row * 2112 + slot * 2 + 1];
          if (row_start + rt >= target || row + 1 >= d.mbh) break;
          row_start += rt; ones_before_row += ro; row++;
        }
        // ones among the first (target - row_start) events of `row`
        const u32 need = target - row_start;
        u32 ones_in_row;
        if (need == rt) {
          ones_in_row = ro;
        } else if (need == 0) {
          ones_in_row = 0;
        } else {
          // lanes take macroblocks of the row; find the MB containing the cut, replay it
          const MbRecord* recs = P.rec1 + d.mb_off + (size_t)row * d.mbw;
          u32 acc_t = 0, acc_o = 0;
          ones_in_row = 0;
          bool found = false;
          for (u32 base = 0; base < d.mbw && !found; base += 32) {
            const u32 mbx = base + lane;
            u32 mt = 0, mo = 0;
            if (mbx < d.mbw) mb_slot_counts(recs[mbx], slot, mt, mo, 0xffffffffu);
            u32 pt = mt, po = mo;  // inclusive scan over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const u32 a = __shfl_up_sync(FULL, pt, o), b = __shfl_up_sync(FULL, po, o);
              if (lane >= o) { pt += a; po += b; }
            }
            const bool here = (acc_t + pt >= need) && (acc_t + pt - mt < need);
            const u32 bal = __ballot_sync(FULL, here);
            if (bal) {
              const int src = __ffs(bal) - 1;
              u32 part_o = 0;
              if (lane == src) {
                u32 tt, oo;
                mb_slot_counts(recs[mbx], slot, tt, oo, need - (acc_t + pt - mt));
                part_o = acc_o + (po - mo) + oo;
              }
              ones_in_row = __shfl_sync(FULL, part_o, src);
              found = true;
            } else {
              acc_t += __shfl_sync(FULL, pt, 31);
              acc_o += __shfl_sync(FULL, po, 31);
            }
          }
          if (!found) ones_in_row = ro;
        }
        const u32 ones_upto = ones_before_row + ones_in_row;
        const u32 seg_ones = ones_upto - consumed_ones;
        const u32 seg_tot = target - done;
        T += seg_tot; O += seg_ones;
        consumed_ones = ones_upto;
        done = target;
        if (done < tot) {  // the next event finds total >= 0xfffe -> halve first
          T = T >> 1;                 // ((s+1)>>1)&0x7fff7fff on the packed word
          O = ((O + 1) >> 1) & 0x7fff;
          boundary += 32767u;
        }
      }
      packed = (T << 16) | O;
    }
    if (lane == 0) s_stats[slot] = packed;
  }
  __syncthreads();
  // should_update / savings per slot (cost.rs:1226-1254) against the DEFAULT probabilities
  for (int slot = threadIdx.x; slot < 1056; slot += blockDim.x) {
    const u32 s = s_stats[slot];
    const int nb = (int)(s & 0xffff), total = (int)(s >> 16);
    const u8 oldp = ZW_TAB(kCoeffProbs)[slot];
    u8 newp = oldp;
    int sav = 0;
    if (total != 0) {
      newp = (u8)(255 - (u32)(nb * 255 / total));
      const u8 up = ZW_TAB(kCoeffUpdateProbs)[slot];
      const int old_cost = nb * (int)bit_cost(1, oldp) + (total - nb) * (int)bit_cost(0, oldp) + (int)bit_cost(0, up);
      const int new_cost = nb * (int)bit_cost(1, newp) + (total - nb) * (int)bit_cost(0, newp) + (int)bit_cost(1, up) + 8 * 256;
      sav = old_cost - new_cost;
    }
    s_newp[slot] = newp;
    s_sav[slot] = sav;
    if (sav > 0) { atomicAdd(&s_total, (u32)sav); atomicAdd(&s_nupd, 1); }
    P.stats[(size_t)img * 1056 + slot] = s;
  }
  __syncthreads();
  const bool use_updated = (i32)s_total > 0 && s_nupd > 0;
  u8* probs = P.probs + (size_t)img * 1056;
  for (int slot = threadIdx.x; slot < 1056; slot += blockDim.x) {
    const u8 p = (use_updated && s_sav[slot] > 0) ? s_newp[slot] : ZW_TAB(kCoeffProbs)[slot];
    s_newp[slot] = p;
    probs[slot] = p;
  }
  __syncthreads();
  u16* lc = P.lcost + (size_t)img * 6528;
  for (int i = threadIdx.x; i < 6528; i += blockDim.x) {
    const int v = i % 68, row = i / 68, ctx = row % 3;
    lc[i] = level_cost_entry(&s_newp[row * 11], ctx, v);
  }
  if (threadIdx.x == 0) {
    const u32 total_mb = d.mbw * d.mbh;
    const u32 non_skip = total_mb - IS.n_skip1;
    u32 prob = (255 * non_skip + total_mb / 2) / total_mb;
    prob = prob > 255 ? 255 : prob;
    IS.skip_prob = (u8)(prob < 1 ? 1 : (prob > 254 ? 254 : prob));
    IS.probs_updated = use_updated;
  }
}

// ---------------------------------------------------------------------------------------------
// (5a) Tokeniser.  Every boolean the reference hands to its two ArithmeticEncoders, as 16-bit
//      (bit << 8 | prob) symbols, generated in parallel per macroblock and laid out in raster order.
// ---------------------------------------------------------------------------------------------
struct TreeCodes {  // path from the root for each value of a tree, MSB first
  u8 len[12];
  u16 code[12];
};
struct TokenTables {
  i8 tree_dct[22], tree_ymode[8], tree_bmode[18], tree_uv[6], tree_seg[6];
  TreeCodes dct, ymode, bmode, uv, seg;
};
__constant__ TokenTables c_tok;

struct CountSink {
  u32 n = 0;
  __device__ __forceinline__ void put(int bit, int prob) { (void)bit; (void)prob; n++; }
};
struct WriteSink {
  Token* p;
  u32 n = 0;
  __device__ __forceinline__ void put(int bit, int prob) { p[n++] = (Token)(((bit ? 1 : 0) << 8) | (prob & 255)); }
};

template <class S>
__device__ __forceinline__ void put_tree(S& s, const i8* tree, const TreeCodes& tc, const u8* probs, int value, int start) {
  int len = tc.len[value];
  const u32 code = tc.code[value];
  int i = 0;
  if (start == 2) { len -= 1; i = 2; }  // skip the "not EOB" branch (write_with_tree_start_index, start_index 2)
  for (int k = len - 1; k >= 0; k--) {
    const int bit = (code >> k) & 1;
    s.put(bit, probs[i >> 1]);
    i = tree[i + bit];
  }
}
template <class S>
__device__ __forceinline__ void put_literal(S& s, int nbits, u32 v) {
  for (int b = nbits - 1; b >= 0; b--) s.put((v >> b) & 1, 128);
}

// encode_coefficients (vp8.rs:798-958) for already-quantised zig-zag levels
template <class S>
__device__ __forceinline__ void block_tokens(S& s, const i16* zz, int plane, int first, int ctx, const u8* probs /*[4][8][3][11]*/) {
  const u8* pp = probs + plane * (8 * 3 * 11);
  int eob = 0;
  for (int i = 0; i < 16; i++) if (zz[i] != 0) eob = i + 1;
  bool skip_eob = false;
  int complexity = ctx;
  for (int idx = first; idx < eob; idx++) {
    const int coeff = zz[idx];
    const int a = iabs(coeff);
    const u8* pr = pp + (ZW_TAB(kCoeffBands)[idx] * 3 + complexity) * 11;
    const int st = skip_eob ? 2 : 0;
    int token;
    if (a == 0) {
      put_tree(s, c_tok.tree_dct, c_tok.dct, pr, 0, st);
      skip_eob = true;
      token = 0;
    } else if (a <= 4) {
      put_tree(s, c_tok.tree_dct, c_tok.dct, pr, a, st);
      skip_eob = false;
      token = a;
    } else {
      int cat;
      if (a <= 6) cat = 5; else if (a <= 10) cat = 6; else if (a <= 18) cat = 7; else if (a <= 34) cat = 8; else if (a <= 66) cat = 9; else cat = 10;
      put_tree(s, c_tok.tree_dct, c_tok.dct, pr, cat, st);
      const u8* cp = &ZW_TAB(kProbDctCat)[(cat - 5) * 12];
      const int extra = a - (int)ZW_TAB(kDctCatBase)[cat - 5];
      int mask = cat == 10 ? (1 << 10) : (1 << (cat - 5));
      for (int k = 0; k < 12; k++) {
        const int prob = cp[k];
        if (prob == 0) break;
        s.put((extra & mask) > 0, prob);
        mask >>= 1;
      }
      skip_eob = false;
      token = cat;
    }
    if (token != 0) s.put(!(coeff > 0), 128);
    complexity = token == 0 ? 0 : (token == 1 ? 1 : 2);
  }
  if (eob < 16) {
    const int bi = eob > first ? eob : first;
    put_tree(s, c_tok.tree_dct, c_tok.dct, pp + (ZW_TAB(kCoeffBands)[bi] * 3 + complexity) * 11, 11, 0);
  }
}

// The intra mode a macroblock leaves as b-pred context (write_macroblock_header :542-551).
__device__ __forceinline__ int ctx_bmode(const MbRecord& r, int sub) {
  if (r.ymode == 4) return r.bmodes[sub];
  // LumaMode::into_intra: DC->DC(0), V->VE(2), H->HE(3), TM->TM(1)
  return r.ymode == 0 ? 0 : (r.ymode == 1 ? 2 : (r.ymode == 2 ? 3 : 1));
}

// write_macroblock_header (vp8.rs:498-560), split into 18 independent pieces so that 18 lanes share the
// (for B_PRED macroblocks long) header instead of one lane walking 16 mode trees alone:
//   slot 0: segment id (if update_map), skip flag, ymode      slots 1..16: sub-block mode i = slot - 1
//   slot 17: uv mode.                                          Concatenated in slot order == reference order.
template <class S>
__device__ __forceinline__ void mb_header_slot(S& s, int slot, const ImageState& IS, const MbRecord& r, const MbRecord* top,
                                               const MbRecord* left) {
  if (slot == 0) {
    if (IS.seg_enabled && IS.update_map) put_tree(s, c_tok.tree_seg, c_tok.seg, IS.tree_probs, r.segment, 0);
    s.put(r.skip, IS.skip_prob);
    put_tree(s, c_tok.tree_ymode, c_tok.ymode, ZW_TAB(kKfYmodeProbs), r.ymode, 0);
  } else if (slot <= 16) {
    if (r.ymode == 4) {
      const int i = slot - 1, x = i & 3, y = i >> 2;
      const int t = y > 0 ? r.bmodes[(y - 1) * 4 + x] : (top ? ctx_bmode(*top, 12 + x) : 0);
      const int l = x > 0 ? r.bmodes[y * 4 + x - 1] : (left ? ctx_bmode(*left, y * 4 + 3) : 0);
      put_tree(s, c_tok.tree_bmode, c_tok.bmode, &ZW_TAB(kKfBmodeProbs)[(t * 10 + l) * 9], r.bmodes[i], 0);
    }
  } else if (slot == 17) {
    put_tree(s, c_tok.tree_uv, c_tok.uv, ZW_TAB(kKfUvModeProbs), r.uvmode, 0);
  }
}
template <class S>
__device__ __forceinline__ void mb_header_tokens(S& s, const ImageState& IS, const MbRecord& r, const MbRecord* top,
                                                 const MbRecord* left) {
  for (int slot = 0; slot < 18; slot++) mb_header_slot(s, slot, IS, r, top, left);
}

// encode_compressed_frame_header (vp8.rs:332-372) incl. segment header (:393-437), quantiser
// indices (:445-458) and the 1056 probability-update flags (:462-496).
template <class S>
__device__ void frame_header_tokens(S& s, const ChunkParams& P, const ImageState& IS, const u8* probs) {
  s.put(0, 128);  // colour space
  s.put(0, 128);  // clamping type
  s.put(IS.seg_enabled, 128);
  if (IS.seg_enabled) {
    s.put(IS.update_map, 128);
    s.put(1, 128);  // update_segment_feature_data
    s.put(0, 128);  // delta mode
    for (int k = 0; k < 4; k++) {
      const int dl = IS.seg_delta[k];
      s.put(dl != 0, 128);
      if (dl != 0) {
        put_literal(s, 7, (u32)iabs(dl));
        s.put(dl < 0, 128);
      }
    }
    for (int k = 0; k < 4; k++) s.put(0, 128);  // loop-filter deltas absent
    if (IS.update_map)
      for (int k = 0; k < 3; k++) {
        const int p = IS.tree_probs[k];
        s.put(p != 255, 128);
        if (p != 255) put_literal(s, 8, (u32)p);
      }
  }
  s.put(0, 128);  // filter_type: normal
  put_literal(s, 6, P.filter_level);
  put_literal(s, 3, 0);  // sharpness
  s.put(0, 128);         // loop_filter_adjustments
  put_literal(s, 2, 0);  // log2(number of token partitions) (always 1 partition, D2)
  put_literal(s, 7, P.base_qidx);
  for (int k = 0; k < 5; k++) s.put(0, 128);  // quantiser deltas absent
  s.put(0, 128);                              // refresh_entropy_probs
  for (int slot = 0; slot < 1056; slot++) {
    const int up = ZW_TAB(kCoeffUpdateProbs)[slot];
    const int np = probs[slot];
    if (np != ZW_TAB(kCoeffProbs)[slot]) {
      s.put(1, up);
      put_literal(s, 8, (u32)np);
    } else {
      s.put(0, up);
    }
  }
  s.put(1, 128);  // mb_no_skip_coeff
  put_literal(s, 8, IS.skip_prob);
}

// MODE 0: count tokens per macroblock.  MODE 1: emit them at the scanned offsets.
// One warp per macroblock: lanes 0..24 = residual blocks; the macroblock header is shared by lanes 0..17.
// MODE 0 also keeps the per-lane counts (block tokens | header-slot tokens << 16), so MODE 1 does not walk the levels a
// second time just to find its offsets.  MODE 1 builds the macroblock's symbols in shared memory (the lanes write 2-byte
// symbols at unrelated offsets) and the warp then copies them out with full 128-byte stores; a macroblock with more
// symbols than the staging buffer holds (rare: > 6 symbols per pixel) is written in place.
#ifndef ZW_TOK_WARPS
#define ZW_TOK_WARPS 2  // macroblocks finish at very different times: small CTAs free their slots early (measured 8 -> 4 -> 2 warps: 4.59 -> 4.28 -> 4.12 ms; 16: 5.56)
#endif
constexpr int TOK_WARPS = ZW_TOK_WARPS;
constexpr u32 TOK_STAGE = 1536;  // symbols per warp (3 KB)
__device__ __forceinline__ void tok_copy_out(const Token* sm, Token* dst, u32 n, int lane) {
  if (n == 0) return;
  const u32 head = (u32)((reinterpret_cast<size_t>(dst) >> 1) & 1u);  // one symbol up to 4-byte alignment
  if (head && lane == 0) dst[0] = sm[0];
  const u32 pairs = (n - (head ? 1u : 0u)) >> 1;
  u32* d32 = reinterpret_cast<u32*>(dst + head);
  for (u32 k = lane; k < pairs; k += 32) d32[k] = (u32)sm[head + 2 * k] | ((u32)sm[head + 2 * k + 1] << 16);
  if (((n - head) & 1u) && lane == 31) dst[n - 1] = sm[n - 1];
}
template <int MODE>
__global__ void __launch_bounds__(TOK_WARPS * 32) k_tokenize(ChunkParams P) {
  __shared__ Token s_stage[MODE == 1 ? TOK_WARPS : 1][MODE == 1 ? TOK_STAGE : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.y;
  const ImageDesc d = P.img[img];
  const ImageState& IS = P.st[img];
  const u32 nmb = d.mbw * d.mbh;
  const u32 mb = blockIdx.x * TOK_WARPS + warp;
  if (mb >= nmb) return;
  if (MODE == 1 && P.tot->overflow) return;  // stream arenas too small: the host grows them and re-runs
  const u32 gmb = d.mb_off + mb;
  const int mbx = mb % d.mbw, mby = mb / d.mbw;
  const MbRecord& r = P.rec2[gmb];
  const u8* probs = P.probs + (size_t)img * 1056;
  const bool nz = lane < 25 && block_nonzero(r.levels[lane < 25 ? lane : 0]);
  const u32 nzmask = __ballot_sync(FULL, nz);
  const BlockInfo I = block_info(r, lane, nzmask);
  const bool do_block = I.coded && !r.skip;
  const MbRecord* top = mby > 0 ? &P.rec2[gmb - d.mbw] : nullptr;
  const MbRecord* left = mbx > 0 ? &P.rec2[gmb - 1] : nullptr;
  if (MODE == 0) {
    u32 cnt = 0;
    if (do_block) {
      CountSink s;
      block_tokens(s, r.levels[lane], I.type, I.first, I.ctx, probs);
      cnt = s.n;
    }
    CountSink hs;
    if (lane < 18) mb_header_slot(hs, lane, IS, r, top, left);
    P.mb_lane_cnt[(size_t)gmb * 32 + lane] = cnt | (hs.n << 16);
    u32 tot = cnt, ht = hs.n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { tot += __shfl_xor_sync(FULL, tot, o); ht += __shfl_xor_sync(FULL, ht, o); }
    if (lane == 0) { P.mb_tok_cnt[gmb] = tot; P.mb_hdr_cnt[gmb] = ht; }
  } else {
    const u32 packed = P.mb_lane_cnt[(size_t)gmb * 32 + lane];
    const u32 cnt = packed & 0xFFFFu, hcnt = packed >> 16;
    u32 incl = cnt, hincl = hcnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 a = __shfl_up_sync(FULL, incl, o), b = __shfl_up_sync(FULL, hincl, o);
      if (lane >= o) { incl += a; hincl += b; }
    }
    const u32 total = __shfl_sync(FULL, incl, 31), htotal = __shfl_sync(FULL, hincl, 31);
    Token* dst = P.tok_tokens + P.lay[img].tok_off + P.mb_tok_cnt[gmb];
    Token* hdst = P.hdr_tokens + P.lay[img].hdr_off + P.mb_hdr_cnt[gmb];  // frame-header length is folded into the scan
    Token* stage = s_stage[warp];
    const bool staged = total <= TOK_STAGE;
    if (do_block) {
      WriteSink s;
      s.p = (staged ? stage : dst) + (incl - cnt);
      block_tokens(s, r.levels[lane], I.type, I.first, I.ctx, probs);
    }
    __syncwarp();
    if (staged) tok_copy_out(stage, dst, total, lane);
    __syncwarp();
    if (lane < 18) {  // the header: at most 122 symbols
      WriteSink s;
      s.p = stage + (hincl - hcnt);
      mb_header_slot(s, lane, IS, r, top, left);
    }
    __syncwarp();
    tok_copy_out(stage, hdst, htotal, lane);
  }
}

// Per image: exclusive scan of the per-macroblock counts (in place) + frame-header token count.
__global__ void __launch_bounds__(256) k_tokscan(ChunkParams P) {
  const int img = blockIdx.x;
  const ImageDesc d = P.img[img];
  ImageState& IS = P.st[img];
  const u32 nmb = d.mbw * d.mbh;
  __shared__ u32 s_part[2][256];
  __shared__ u32 s_fh;
  if (threadIdx.x == 0) {
    CountSink s;
    frame_header_tokens(s, P, IS, P.probs + (size_t)img * 1056);
    s_fh = s.n;
  }
  const u32 per = (nmb + blockDim.x - 1) / blockDim.x;
  const u32 b0 = min(threadIdx.x * per, nmb), b1 = min(b0 + per, nmb);
  u32 sh = 0, st = 0;
  for (u32 i = b0; i < b1; i++) { sh += P.mb_hdr_cnt[d.mb_off + i]; st += P.mb_tok_cnt[d.mb_off + i]; }
  s_part[0][threadIdx.x] = sh;
  s_part[1][threadIdx.x] = st;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 ah = s_fh, at = 0;
    for (int i = 0; i < (int)blockDim.x; i++) {
      const u32 h = s_part[0][i], t = s_part[1][i];
      s_part[0][i] = ah; s_part[1][i] = at;
      ah += h; at += t;
    }
    IS.hdr_tokens = ah;
    IS.tok_tokens = at;
  }
  __syncthreads();
  u32 ah = s_part[0][threadIdx.x], at = s_part[1][threadIdx.x];
  for (u32 i = b0; i < b1; i++) {
    const u32 h = P.mb_hdr_cnt[d.mb_off + i], t = P.mb_tok_cnt[d.mb_off + i];
    P.mb_hdr_cnt[d.mb_off + i] = ah;
    P.mb_tok_cnt[d.mb_off + i] = at;
    ah += h; at += t;
  }
}

// Chunk-wide placement of the variable-size data (one CTA): exclusive scans over the images of the symbol counts
// (8-token aligned), of the partition capacities (7 bits per symbol bound + 16, 16-byte aligned pairs) and of the
// boolean-coder segment counts, checked against the arenas the host allocated before it knew the counts.
__global__ void __launch_bounds__(1024) k_layout(ChunkParams P) {
  constexpr int NQ = 5;  // scanned quantities: hdr tokens, tok tokens, partition bytes, tok segments, hdr segments
  __shared__ u64 s_w[NQ][32];
  __shared__ u64 s_base[NQ];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NQ) s_base[threadIdx.x] = 0;
  __syncthreads();
  for (u32 i0 = 0; i0 < P.n_img; i0 += blockDim.x) {
    const u32 i = i0 + threadIdx.x;
    u64 v[NQ] = {0, 0, 0, 0, 0}, inc[NQ];
    u32 c0 = 0, c1 = 0;
    if (i < P.n_img) {
      const ImageState& IS = P.st[i];
      v[0] = ((u64)IS.hdr_tokens + 7) & ~7ull;
      v[1] = ((u64)IS.tok_tokens + 7) & ~7ull;
      c0 = (u32)(((u64)IS.hdr_tokens * 7) / 8 + 16);
      c1 = (u32)(((u64)IS.tok_tokens * 7) / 8 + 16);
      v[2] = ((u64)c0 + c1 + 15) & ~15ull;
      v[3] = bc_segments(IS.tok_tokens);
      v[4] = bc_segments(IS.hdr_tokens);
    }
#pragma unroll
    for (int q = 0; q < NQ; q++) {  // inclusive scans inside the warp
      u64 x = v[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 y = (u64)shfl_up64_full((i64)x, o);
        if (lane >= o) x += y;
      }
      inc[q] = x;
      if (lane == 31) s_w[q][warp] = x;
    }
    __syncthreads();
    if (warp < NQ) {  // warp q scans the 32 warp totals of quantity q (exclusive)
      const u64 w0 = s_w[warp][lane];
      u64 x = w0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 y = (u64)shfl_up64_full((i64)x, o);
        if (lane >= o) x += y;
      }
      s_w[warp][lane] = x - w0;
    }
    __syncthreads();
    u64 ex[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) ex[q] = s_base[q] + s_w[q][warp] + inc[q] - v[q];
    if (i < P.n_img) {
      ImageLayout L;
      L.hdr_off = ex[0]; L.tok_off = ex[1]; L.part_off = ex[2];
      L.out_off = 0;
      L.p0_cap = c0; L.p1_cap = c1;
      P.lay[i] = L;
      P.seg_off[i] = (u32)ex[3];            // token-partition streams first ...
      P.seg_off[P.n_img + i] = (u32)ex[4];  // ... first partitions after them (offset added below)
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) {  // the last thread holds the totals of this step
#pragma unroll
      for (int q = 0; q < NQ; q++) s_base[q] = ex[q] + v[q];
    }
    __syncthreads();
  }
  const u32 tseg = (u32)s_base[3], hseg = (u32)s_base[4];
  for (u32 i = threadIdx.x; i < P.n_img; i += blockDim.x) P.seg_off[P.n_img + i] += tseg;
  if (threadIdx.x == 0) {
    P.seg_off[2 * P.n_img] = tseg + hseg;
    ChunkTotals T;
    T.hdr_tokens = s_base[0]; T.tok_tokens = s_base[1]; T.part_bytes = s_base[2]; T.out_bytes = 0;
    T.overflow = (T.hdr_tokens > P.cap_hdr_tokens || T.tok_tokens > P.cap_tok_tokens || T.part_bytes > P.cap_part_bytes ||
                  (u64)tseg + hseg > P.cap_segments) ? 1u : 0u;
    T.segments = tseg + hseg;
    *P.tot = T;
  }
}

// Frame-header tokens of every image (tiny, one thread each).
__global__ void k_frame_header(ChunkParams P) {
  const u32 img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= P.n_img || P.tot->overflow) return;
  WriteSink s;
  s.p = P.hdr_tokens + P.lay[img].hdr_off;
  frame_header_tokens(s, P, P.st[img], P.probs + (size_t)img * 1056);
}

// ---------------------------------------------------------------------------------------------
// (5b) Boolean entropy coder (ArithmeticEncoder, arithmetic.rs:7-196), segment-parallel and exact.
//
// write_bool is a serial recurrence per stream, but its two halves separate:
//   * the RANGE after a symbol depends only on the range before it and the symbol; renormalised it takes one of
//     128 values (range - 1 in [127, 254]);
//   * `bottom` with its carries is a sum: symbol k adds its 8-bit split at stream bit P_k = s_0 + .. + s_(k-1)
//     (s = renormalisation shifts).  Sums can be formed piecewise and added up afterwards.
// So a stream is cut into segments of BC_SEG symbols and every segment is coded by ONE LANE with the
// reference's own serial algorithm, all segments of all streams of the batch in parallel, once each segment knows
// the range it starts with and its start bit:
//   k_bc_cands   which range states can a segment possibly start in?  All 128 states are run through the BC_WARM
//                symbols before the segment start; the state map is many-to-one, so a handful (1..6) survive.
//   k_bc_trans   per segment and surviving start state: end state and total shift of the segment (one lane each).
//   k_bc_resolve per stream: walk the segments, picking the transition of the state actually reached -> start state
//                and start bit of every segment.  (The true state at the warm-up start is one of the 128, so the
//                true segment start state is always among the candidates: no speculation, no fallback.)
//   k_bc_code    one lane per segment: the reference's write_bool from (state, bottom = 0, bit position), bytes
//                written straight to their final place; what is still pending in `bottom` at the segment end
//                (the tail) and carries that leave the segment's own bytes are recorded.
//   k_bc_fix     per stream: add every tail to the bytes that follow it, ripple the recorded carries
//                (add_one_to_output, arithmetic.rs:47-60), set the partition sizes.
// Round 1 ran one warp PAIR per stream with all 32 lanes of the chain warp on the same recurrence: 47 ms for the
// 1024-photo batch (2.4 symbols per pixel), bound by the per-symbol latency of ~1000 concurrent chains.
// ---------------------------------------------------------------------------------------------
struct BcStream {
  const Token* tk;
  u32 n;     // symbols
  u8* out;   // coded partition
  u32 cap;
  u32 img;
  bool is_hdr;
};
__device__ __forceinline__ BcStream bc_stream(const ChunkParams& P, u32 sid) {
  BcStream S;
  S.is_hdr = sid >= P.n_img;
  S.img = S.is_hdr ? sid - P.n_img : sid;
  const ImageLayout L = P.lay[S.img];
  const ImageState& IS = P.st[S.img];
  S.tk = S.is_hdr ? P.hdr_tokens + L.hdr_off : P.tok_tokens + L.tok_off;
  S.n = S.is_hdr ? IS.hdr_tokens : IS.tok_tokens;
  S.out = P.part_bytes + L.part_off + (S.is_hdr ? 0 : L.p0_cap);  // partition scratch: [first | token]
  S.cap = S.is_hdr ? L.p0_cap : L.p1_cap;
  return S;
}
// global segment index -> (stream, segment within the stream)
__device__ __forceinline__ void bc_locate(const ChunkParams& P, u32 g, u32& sid, u32& j) {
  u32 lo = 0, hi = 2 * P.n_img;  // seg_off[lo] <= g < seg_off[hi]
  while (hi - lo > 1) {
    const u32 mid = (lo + hi) >> 1;
    if (P.seg_off[mid] <= g) lo = mid; else hi = mid;
  }
  sid = lo;
  j = g - P.seg_off[lo];
}
constexpr int BC_CAND_FIRST = 64;  // symbols all 128 states are run through before the survivors move into one warp
__global__ void __launch_bounds__(128) k_bc_cands(ChunkParams P) {
  if (P.tot->overflow) return;
  __shared__ u32 s_map[4];
  const u32 total = P.seg_off[2 * P.n_img];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (u32 g = blockIdx.x; g < total; g += gridDim.x) {
    u32 sid, j;
    bc_locate(P, g, sid, j);
    if (j == 0) {  // a stream starts with range 255
      if (tid < 4) P.seg[g].cand[tid] = tid == 3 ? 0x80000000u : 0u;
      continue;
    }
    const BcStream S = bc_stream(P, sid);
    const Token* tk = S.tk + (size_t)j * BC_SEG - BC_WARM;  // 16-byte aligned: stream starts and BC_SEG, BC_WARM are multiples of 8 tokens
    u32 st = 127u + (u32)tid, add, sh;
    auto run = [&](u32 from, u32 to) {
      uint4 qn = __ldg(reinterpret_cast<const uint4*>(tk + from));  // the same address in every lane: one broadcast; one group ahead
      for (u32 i = from; i < to; i += 8) {
        const uint4 q = qn;
        if (i + 8 < to) qn = __ldg(reinterpret_cast<const uint4*>(tk + i + 8));
#pragma unroll
        for (int e = 0; e < 8; e++) st = bc_step(st, bc_sym8(q.x, q.y, q.z, q.w, e), add, sh);
      }
    };
    run(0, BC_CAND_FIRST);
    if (tid < 4) s_map[tid] = 0;
    __syncthreads();
    atomicOr(&s_map[(st - 127u) >> 5], 1u << ((st - 127u) & 31u));
    __syncthreads();
    const u32 m0 = s_map[0], m1 = s_map[1], m2 = s_map[2], m3 = s_map[3];
    const u32 k = (u32)(__popc(m0) + __popc(m1) + __popc(m2) + __popc(m3));
    const bool narrow = k <= 32;  // the survivors fit one warp (always, on real streams): three warps retire
    if (narrow) {
      if (warp == 0) {
        st = 127u + bc_nth_bit(m0, m1, m2, m3, (u32)tid < k ? (u32)tid : 0u);
        run(BC_CAND_FIRST, BC_WARM);
      }
    } else {
      run(BC_CAND_FIRST, BC_WARM);
    }
    __syncthreads();
    if (tid < 4) s_map[tid] = 0;
    __syncthreads();
    if (!narrow || warp == 0) atomicOr(&s_map[(st - 127u) >> 5], 1u << ((st - 127u) & 31u));
    __syncthreads();
    if (tid < 4) P.seg[g].cand[tid] = s_map[tid];
    __syncthreads();
  }
}

// Four lanes per segment: lane q of a group takes the candidate start states of rank q, q + 4, ...
__global__ void __launch_bounds__(128) k_bc_trans(ChunkParams P) {
  if (P.tot->overflow) return;
  const u32 total = P.seg_off[2 * P.n_img];
  const u32 gl = threadIdx.x & 3u;
  const u32 stride = (gridDim.x * blockDim.x) >> 2;
  for (u32 g = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; g < total; g += stride) {
    u32 sid, j;
    bc_locate(P, g, sid, j);
    const BcStream S = bc_stream(P, sid);
    const u32 first = j * BC_SEG;
    const u32 n = S.n > first ? min(BC_SEG, S.n - first) : 0u;
    const Token* tk = S.tk + first;
    const u32 m0 = P.seg[g].cand[0], m1 = P.seg[g].cand[1], m2 = P.seg[g].cand[2], m3 = P.seg[g].cand[3];
    const u32 k = (u32)(__popc(m0) + __popc(m1) + __popc(m2) + __popc(m3));
    for (u32 c = gl; c < k; c += 4) {
      u32 st = 127u + bc_nth_bit(m0, m1, m2, m3, c), T = 0, add, sh;
      u32 i = 0;
      // the symbols are requested two groups of eight ahead of the group being walked (a lane's chain otherwise pays one full
      // memory latency per 16-byte load: the load sits behind the dependent steps of the previous group)
      uint4 qn = n >= 8 ? __ldg(reinterpret_cast<const uint4*>(tk)) : make_uint4(0, 0, 0, 0);
      uint4 qm = n >= 16 ? __ldg(reinterpret_cast<const uint4*>(tk + 8)) : make_uint4(0, 0, 0, 0);
      for (; i + 8 <= n; i += 8) {
        const uint4 q = qn;
        qn = qm;
        if (i + 24 <= n) qm = __ldg(reinterpret_cast<const uint4*>(tk + i + 16));
#pragma unroll
        for (int e = 0; e < 8; e++) { st = bc_step(st, bc_sym8(q.x, q.y, q.z, q.w, e), add, sh); T += sh; }
      }
      for (; i < n; i++) { st = bc_step(st, tk[i], add, sh); T += sh; }
      P.seg_trans[(size_t)g * 128 + c] = st | (T << 8);
    }
  }
}

// One thread per stream: the state and the bit position every segment starts with.
__global__ void __launch_bounds__(128) k_bc_resolve(ChunkParams P) {
  if (P.tot->overflow) return;
  const u32 sid = blockIdx.x * blockDim.x + threadIdx.x;
  if (sid >= 2 * P.n_img) return;
  const u32 g0 = P.seg_off[sid], g1 = P.seg_off[sid + 1];
  u32 st = 254;
  u64 bits = 0;
  for (u32 g = g0; g < g1; g++) {
    BcSegment& sg = P.seg[g];
    const u32 e = P.seg_trans[(size_t)g * 128 + bc_rank(sg.cand[0], sg.cand[1], sg.cand[2], sg.cand[3], st - 127u)];
    sg.state = (u8)st;
    sg.start_bit = bits;
    st = e & 255u;
    bits += e >> 8;
  }
}

// One LANE per segment: write_bool (arithmetic.rs:67-95) with the bit-at-a-time renormalisation loop collapsed into at
// most two steps around the byte boundary, starting from bottom = 0.
__global__ void __launch_bounds__(128) k_bc_code(ChunkParams P) {
  if (P.tot->overflow) return;
  const u32 total = P.seg_off[2 * P.n_img];
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
    u32 sid, j;
    bc_locate(P, g, sid, j);
    const BcStream S = bc_stream(P, sid);
    const bool last = g + 1 == P.seg_off[sid + 1];
    const u32 first = j * BC_SEG;
    const u32 n = S.n > first ? min(BC_SEG, S.n - first) : 0u;
    const Token* tk = S.tk + first;
    BcSegment& sg = P.seg[g];
    BcCoder cd;
    cd.begin(sg.state, sg.start_bit, S.out, S.cap);
    u32 i = 0;
    uint4 qn = n >= 8 ? __ldg(reinterpret_cast<const uint4*>(tk)) : make_uint4(0, 0, 0, 0);
    uint4 qm = n >= 16 ? __ldg(reinterpret_cast<const uint4*>(tk + 8)) : make_uint4(0, 0, 0, 0);
    for (; i + 8 <= n; i += 8) {
      const uint4 q = qn;
      qn = qm;
      if (i + 24 <= n) qm = __ldg(reinterpret_cast<const uint4*>(tk + i + 16));
#pragma unroll
      for (int e = 0; e < 8; e++) cd.put(bc_sym8(q.x, q.y, q.z, q.w, e));
    }
    for (; i < n; i++) cd.put(tk[i]);
    if (!last) {
      sg.tail = cd.tail();
    } else {
      const u32 bytes = cd.flush();
      sg.tail = 0;
      ImageState& IS = P.st[S.img];
      if (S.is_hdr) IS.part0_bytes = bytes; else IS.part1_bytes = bytes;
    }
    const u32 carries = cd.carries;
    const bool overflow = cd.overflow;
    sg.carries = carries;
    if (overflow) P.st[S.img].status = 4;  // ZW_ERR_OUTPUT_TOO_SMALL (cannot happen with the 7-bits-per-symbol bound)
  }
}

// One thread per stream: add the tail of every segment to the bytes that follow it and ripple the carries.
__global__ void __launch_bounds__(128) k_bc_fix(ChunkParams P) {
  if (P.tot->overflow) return;
  const u32 sid = blockIdx.x * blockDim.x + threadIdx.x;
  if (sid >= 2 * P.n_img) return;
  const BcStream S = bc_stream(P, sid);
  if (P.st[S.img].status == 4) return;
  const u32 g0 = P.seg_off[sid], g1 = P.seg_off[sid + 1];
  for (u32 g = g0 + 1; g < g1; g++) {
    bc_fix_boundary(S.out, P.seg[g].start_bit, P.seg[g - 1].tail, P.seg[g].carries);
  }
}

// ---------------------------------------------------------------------------------------------
// (6) Assembly: frame tag + start code + dimensions + first partition + token partition, packed
//     back to back in the output arena in image order.  (Only the RIFF wrap stays on the host.)
// ---------------------------------------------------------------------------------------------
__global__ void k_outscan(ChunkParams P, u64* out_offsets /*[n_img+1]*/) {
  // one warp: exclusive scan of the 16-byte-aligned file sizes, 32 images per step
  if (blockIdx.x != 0 || threadIdx.x >= 32 || P.tot->overflow) return;
  const int lane = threadIdx.x;
  u64 acc = 0;
  for (u32 i0 = 0; i0 < P.n_img; i0 += 32) {
    const u32 i = i0 + lane;
    u64 sz = 0;
    if (i < P.n_img) {
      ImageState& IS = P.st[i];
      IS.vp8_bytes = 10 + IS.part0_bytes + IS.part1_bytes;
      IS.file_bytes = 20 + IS.vp8_bytes + (IS.vp8_bytes & 1);
      if (IS.part0_bytes >= (1u << 19) && IS.status == 0) IS.status = 5;  // ZW_ERR_PARTITION_TOO_LARGE
      sz = (IS.file_bytes + 15u) & ~15u;
    }
    u64 incl = sz;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 v = (u64)shfl_up64_full((i64)incl, o);
      if (lane >= o) incl += v;
    }
    if (i < P.n_img) { out_offsets[i] = acc + incl - sz; P.lay[i].out_off = acc + incl - sz; }
    acc += (u64)shfl64((i64)incl, 31);
  }
  if (lane == 0) { out_offsets[P.n_img] = acc; P.tot->out_bytes = acc; }
}

// The finished file of every image: RIFF header (api.rs:1325-1329, write_chunk :1232-1241) + frame tag + start code +
// dimensions + first partition + token partition + pad byte.  The VP8 payload alone is bytes [20, 20 + vp8_bytes).
__global__ void __launch_bounds__(256) k_assemble(ChunkParams P, const u64* out_offsets) {
  if (P.tot->overflow) return;
  const int img = blockIdx.x;
  const ImageDesc d = P.img[img];
  const ImageLayout L = P.lay[img];
  const ImageState& IS = P.st[img];
  u8* o = P.out + out_offsets[img];
  const u8* p0 = P.part_bytes + L.part_off;
  const u8* p1 = p0 + L.p0_cap;
  if (threadIdx.x == 0) {
    const u32 payload = IS.vp8_bytes;
    const u32 chunk = payload + (payload & 1) + 8, riff = chunk + 4;
    o[0] = 'R'; o[1] = 'I'; o[2] = 'F'; o[3] = 'F';
    o[4] = (u8)riff; o[5] = (u8)(riff >> 8); o[6] = (u8)(riff >> 16); o[7] = (u8)(riff >> 24);
    o[8] = 'W'; o[9] = 'E'; o[10] = 'B'; o[11] = 'P'; o[12] = 'V'; o[13] = 'P'; o[14] = '8'; o[15] = ' ';
    o[16] = (u8)payload; o[17] = (u8)(payload >> 8); o[18] = (u8)(payload >> 16); o[19] = (u8)(payload >> 24);
    if (payload & 1) o[20 + payload] = 0;
    u8* v = o + 20;
    const u32 tag = (IS.part0_bytes << 5) | (1u << 4);  // show_frame=1, version 0, key frame
    v[0] = (u8)tag; v[1] = (u8)(tag >> 8); v[2] = (u8)(tag >> 16);
    v[3] = 0x9d; v[4] = 0x01; v[5] = 0x2a;
    const u32 w = d.width & 0x3fff, h = d.height & 0x3fff;
    v[6] = (u8)w; v[7] = (u8)(w >> 8); v[8] = (u8)h; v[9] = (u8)(h >> 8);
  }
  for (u32 i = threadIdx.x; i < IS.part0_bytes; i += blockDim.x) o[30 + i] = p0[i];
  u8* o1 = o + 30 + IS.part0_bytes;
  for (u32 i = threadIdx.x; i < IS.part1_bytes; i += blockDim.x) o1[i] = p1[i];
}

}  // namespace zw
#endif
