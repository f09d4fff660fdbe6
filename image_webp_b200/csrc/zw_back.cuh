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
//   ArithmeticEncoder       src/encoder/arithmetic.rs:7-196    (k_boolcode)
//   write_uncompressed_frame_header / write_partitions vp8.rs:315-330, :374-391 (k_assemble)
#ifndef ZW_BACK_CUH
#define ZW_BACK_CUH
#include "zw_search.cuh"

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
constexpr int STAT_WARPS = 4;

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
constexpr int TOK_WARPS = 8;
template <int MODE>
__global__ void __launch_bounds__(TOK_WARPS * 32) k_tokenize(ChunkParams P) {
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
  u32 cnt = 0;
  if (MODE == 0) {
    if (do_block) {
      CountSink s;
      block_tokens(s, r.levels[lane], I.type, I.first, I.ctx, probs);
      cnt = s.n;
    }
    u32 tot = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
    if (lane == 0) P.mb_tok_cnt[gmb] = tot;
    {
      CountSink s;
      if (lane < 18) mb_header_slot(s, lane, IS, r, mby > 0 ? &P.rec2[gmb - d.mbw] : nullptr, mbx > 0 ? &P.rec2[gmb - 1] : nullptr);
      u32 ht = s.n;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ht += __shfl_xor_sync(FULL, ht, o);
      if (lane == 0) P.mb_hdr_cnt[gmb] = ht;
    }
  } else {
    // recount to get the per-block offsets inside the macroblock (cheap; avoids a per-block array)
    if (do_block) {
      CountSink s;
      block_tokens(s, r.levels[lane], I.type, I.first, I.ctx, probs);
      cnt = s.n;
    }
    u32 incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 a = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += a;
    }
    if (do_block) {
      WriteSink s;
      s.p = P.tok_tokens + P.lay[img].tok_off + P.mb_tok_cnt[gmb] + (incl - cnt);
      block_tokens(s, r.levels[lane], I.type, I.first, I.ctx, probs);
    }
    {
      const MbRecord* top = mby > 0 ? &P.rec2[gmb - d.mbw] : nullptr;
      const MbRecord* left = mbx > 0 ? &P.rec2[gmb - 1] : nullptr;
      CountSink c;
      if (lane < 18) mb_header_slot(c, lane, IS, r, top, left);
      u32 hincl = c.n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 a = __shfl_up_sync(FULL, hincl, o);
        if (lane >= o) hincl += a;
      }
      if (lane < 18) {
        WriteSink s;
        s.p = P.hdr_tokens + P.lay[img].hdr_off + P.mb_hdr_cnt[gmb] + (hincl - c.n);  // frame-header length is folded into the scan
        mb_header_slot(s, lane, IS, r, top, left);
      }
    }
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
// (8-token aligned) and of the partition capacities (7 bits per symbol bound + 16, 16-byte aligned pairs), checked
// against the arenas the host allocated before it knew the counts.
__global__ void __launch_bounds__(1024) k_layout(ChunkParams P) {
  __shared__ u64 s_h[32], s_t[32], s_p[32];
  __shared__ u64 s_base[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 3) s_base[threadIdx.x] = 0;
  __syncthreads();
  for (u32 i0 = 0; i0 < P.n_img; i0 += blockDim.x) {
    const u32 i = i0 + threadIdx.x;
    u64 h = 0, t = 0, pb = 0;
    u32 c0 = 0, c1 = 0;
    if (i < P.n_img) {
      const ImageState& IS = P.st[i];
      h = ((u64)IS.hdr_tokens + 7) & ~7ull;
      t = ((u64)IS.tok_tokens + 7) & ~7ull;
      c0 = (u32)(((u64)IS.hdr_tokens * 7) / 8 + 16);
      c1 = (u32)(((u64)IS.tok_tokens * 7) / 8 + 16);
      pb = ((u64)c0 + c1 + 15) & ~15ull;
    }
    u64 ih = h, it = t, ip = pb;  // inclusive scans inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 a = (u64)shfl_up64_full((i64)ih, o), b = (u64)shfl_up64_full((i64)it, o), c = (u64)shfl_up64_full((i64)ip, o);
      if (lane >= o) { ih += a; it += b; ip += c; }
    }
    if (lane == 31) { s_h[warp] = ih; s_t[warp] = it; s_p[warp] = ip; }
    __syncthreads();
    if (warp == 0) {
      u64 a = s_h[lane], b = s_t[lane], c = s_p[lane];
      const u64 a0 = a, b0 = b, c0w = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 x = (u64)shfl_up64_full((i64)a, o), y = (u64)shfl_up64_full((i64)b, o), z = (u64)shfl_up64_full((i64)c, o);
        if (lane >= o) { a += x; b += y; c += z; }
      }
      s_h[lane] = a - a0; s_t[lane] = b - b0; s_p[lane] = c - c0w;  // exclusive warp offsets
    }
    __syncthreads();
    if (i < P.n_img) {
      ImageLayout L;
      L.hdr_off = s_base[0] + s_h[warp] + ih - h;
      L.tok_off = s_base[1] + s_t[warp] + it - t;
      L.part_off = s_base[2] + s_p[warp] + ip - pb;
      L.out_off = 0;
      L.p0_cap = c0; L.p1_cap = c1;
      P.lay[i] = L;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) {  // last thread holds the block totals of this step
      s_base[0] += s_h[warp] + ih; s_base[1] += s_t[warp] + it; s_base[2] += s_p[warp] + ip;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ChunkTotals T;
    T.hdr_tokens = s_base[0]; T.tok_tokens = s_base[1]; T.part_bytes = s_base[2]; T.out_bytes = 0;
    T.overflow = (T.hdr_tokens > P.cap_hdr_tokens || T.tok_tokens > P.cap_tok_tokens || T.part_bytes > P.cap_part_bytes) ? 1u : 0u;
    T.pad = 0;
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
// (5b) Boolean entropy coder: one lane per (image, partition) stream.
//      Streams 0..n_img-1 are the token partitions, n_img..2n_img-1 the first partitions, so the
//      lanes of a warp carry streams of similar length.
// ---------------------------------------------------------------------------------------------
// add_one_to_output (arithmetic.rs:47-60): ripple a carry through trailing 0xFF bytes.
__device__ __forceinline__ void bool_carry(u8* out, u32 pos, u32 cap) {
  u32 j = pos < cap ? pos : cap;
  while (j > 0) {
    j--;
    if (out[j] < 255) { out[j]++; break; }
    out[j] = 0;
  }
}

// One WARP per (image, partition) stream.  Streams 0..n_img-1 are the token partitions,
// n_img..2n_img-1 the first partitions.
//
// The coder is split into its strictly serial part and a parallel part.  write_bool
// (arithmetic.rs:67-95) does, per symbol, `bottom += bit ? split : 0; bottom <<= s` with
// split = 1 + (((range - 1) * prob) >> 8) and s = the renormalisation shift of the new range:
//   * (range, split, s) depend only on the previous range and the symbol -- a short dependent
//     chain (multiply, shift, select, count-leading-zeros, shift) that all lanes run in lock step
//     over the 32 symbols the warp fetched with one coalesced load;
//   * `bottom` with its carries is a long binary number: symbol k contributes its 8-bit `split`
//     with the most significant bit at stream bit P_k = s_0 + .. + s_(k-1) (byte 0 = stream bits
//     0..7: the initial bit_num = 24 aligns the first addend with the first byte).  Lane k adds its
//     symbol's contribution into a shared-memory window of 16-bit cells, the cells are carry-
//     normalised with a ballot-based carry look-ahead, completed cells are flushed as bytes and the
//     window slides.  A carry out of the window ripples into bytes already written exactly like
//     add_one_to_output (arithmetic.rs:47-60).
// The flush (arithmetic.rs:176-195) emits the pending bits and pads to pos + 4 bytes, where
// pos = 0 if T < 24 else 1 + (T - 24) / 8 for T total shifts: the first pos + 4 bytes of the number.
// Two warps per stream, software-pipelined over the groups of 32 symbols: the CHAIN warp runs the range
// recurrence of group g + 1 (parking the range every symbol starts from in a double-buffered shared
// slot) while the SUM warp does the parallel part of group g; one named barrier per group and pair.
constexpr int BC_STREAMS = 2;            // streams per CTA
constexpr int BC_WARPS = 2 * BC_STREAMS;  // warp 2s: chain, warp 2s + 1: sum
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(BC_WARPS * 32) k_boolcode(ChunkParams P) {
  __shared__ u32 s_cells[BC_STREAMS][32];
  __shared__ u32 s_range[BC_STREAMS][2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pair = warp >> 1;
  const bool chain_warp = (warp & 1) == 0;
  const u32 sid = blockIdx.x * BC_STREAMS + pair;
  if (sid >= 2 * P.n_img || P.tot->overflow) return;  // both warps of a pair leave together
  const bool is_hdr = sid >= P.n_img;
  const u32 img = is_hdr ? sid - P.n_img : sid;
  const ImageLayout d = P.lay[img];
  ImageState& IS = P.st[img];
  const Token* tk = is_hdr ? P.hdr_tokens + d.hdr_off : P.tok_tokens + d.tok_off;
  const u32 n = is_hdr ? IS.hdr_tokens : IS.tok_tokens;
  const u32 groups = (n + 31) / 32;
  if (chain_warp) {
    // ---- serial part: the range recurrence; lane 0 parks range - 1 of every symbol ----
    u32 rm1 = 254;  // range - 1, uniform across the warp (the recurrence is shortest in this form)
    u32 mine = lane < n ? tk[lane] : 0;
    for (u32 g = 0; g <= groups; g++) {
      if (g < groups) {
        const u32 i0 = g * 32;
        const u32 nxt = (i0 + 32 + lane < n) ? tk[i0 + 32 + lane] : 0;  // prefetch the next 32 symbols
        const int cnt = (int)(n - i0 < 32 ? n - i0 : 32);
        const u32 mine24 = ((mine & 255u) << 24) | (mine >> 8);  // prob in the top byte, bit in bit 0
        u32* rsave = s_range[pair][g & 1];
        if (cnt == 32) {
#pragma unroll
          for (int k = 0; k < 32; k++) {
            const u32 t = __shfl_sync(FULL, mine24, k);
            if (lane == 0) rsave[k] = rm1;
            const u32 x = __umulhi(rm1, t & 0xff000000u);  // ((range - 1) * prob) >> 8 = split - 1
            const u32 r2 = (t & 1) ? rm1 - x : x + 1;
            const u32 r2m1 = (t & 1) ? rm1 - x - 1 : x;
            rm1 = __funnelshift_l(0xffffffffu, r2m1, __clz(r2) - 24);  // (r2 << s) - 1
          }
        } else {
#pragma unroll 1
          for (int k = 0; k < cnt; k++) {
            const u32 t = __shfl_sync(FULL, mine24, k);
            if (lane == 0) rsave[k] = rm1;
            const u32 x = __umulhi(rm1, t & 0xff000000u);
            const u32 r2 = (t & 1) ? rm1 - x : x + 1;
            const u32 r2m1 = (t & 1) ? rm1 - x - 1 : x;
            rm1 = __funnelshift_l(0xffffffffu, r2m1, __clz(r2) - 24);
          }
        }
        mine = nxt;
      }
      pair_barrier(1 + pair);  // barrier g: slot g & 1 is complete and slot (g + 1) & 1 has been consumed
    }
    return;
  }
  // ---- SUM warp: the parallel part, one group behind the chain ----
  u8* out = P.part_bytes + d.part_off + (is_hdr ? 0 : d.p0_cap);  // partition scratch: [first | token]
  const u32 cap = is_hdr ? d.p0_cap : d.p1_cap;
  u32* cells = s_cells[pair];
  cells[lane] = 0;
  __syncwarp();
  u32 T = 0;        // total renormalisation shifts so far == stream bit of the next addend's MSB
  u32 base = 0;     // stream bit of cell 0 (multiple of 16); base / 8 bytes are already written
  bool overflow = false;
  u32 mine = lane < n ? tk[lane] : 0;
  for (u32 g = 0; g <= groups; g++) {
    pair_barrier(1 + pair);  // barrier g: slot g & 1 is complete; the chain now fills slot (g + 1) & 1
    if (g == groups) break;  // the chain's last barrier
    const u32 i0 = g * 32;
    const u32 nxt = (i0 + 32 + lane < n) ? tk[i0 + 32 + lane] : 0;
    const int cnt = (int)(n - i0 < 32 ? n - i0 : 32);
    const u32* rsave = s_range[pair][g & 1];
    // lane k redoes symbol k from its parked range
    u32 add = 0, sh = 0;
    if (lane < cnt) {
      const u32 r1 = rsave[lane];  // range - 1 this symbol starts from
      const u32 x = (r1 * (mine & 255)) >> 8;
      const bool bit = (mine >> 8) != 0;
      const u32 r2 = bit ? r1 - x : x + 1;
      add = bit ? x + 1 : 0;
      sh = (u32)(__clz(r2) - 24);
    }
    u32 incl = sh;  // inclusive prefix sum of the shifts
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const u32 v = __shfl_up_sync(FULL, incl, dlt);
      if (lane >= dlt) incl += v;
    }
    const u32 total = __shfl_sync(FULL, incl, 31);
    if (add) {
      const u32 o = T + incl - sh - base;  // window offset of the addend's MSB
      const u32 c = o >> 4, b = o & 15;
      const u32 f = add << (24 - b);       // two-cell field: cell c in the high half, c + 1 in the low half
      atomicAdd(&cells[c], f >> 16);
      if (f & 0xffffu) atomicAdd(&cells[c + 1], f & 0xffffu);
    }
    T += total;
    __syncwarp();
    // ---- carry normalisation: cell c keeps 16 bits, the excess moves to cell c - 1 ----
    const u32 v = cells[lane];
    const u32 from_next = __shfl_down_sync(FULL, v >> 16, 1);
    const u32 v1 = (v & 0xffffu) + (lane < 31 ? from_next : 0u);  // <= 0xffff + 31
    // look-ahead on the reversed cell order (bit i <-> cell 31 - i) so that carries move up
    const u32 G = __brev(__ballot_sync(FULL, v1 > 0xffffu)), Pm = __brev(__ballot_sync(FULL, v1 == 0xffffu));
    const u64 sum = (u64)Pm + ((u64)G << 1);
    const u32 cin = ((u32)sum ^ Pm);  // carry into reversed position i
    const u32 my_cin = (cin >> (31 - lane)) & 1;
    const u32 fin = (v1 + my_cin) & 0xffffu;
    // carry out of cell 0 (reversed position 31): generated there or propagated through it
    u32 carry0 = (__shfl_sync(FULL, v, 0) >> 16) + (u32)((sum >> 32) & 1);
    if (lane == 0) {
      for (; carry0 > 0; carry0--) bool_carry(out, base >> 3, cap);
    }
    // ---- flush the cells no later symbol can reach (only carries can, through bool_carry) ----
    const u32 nf = (T - base) >> 4;
    if ((u32)lane < nf) {
      const u32 bp = (base >> 3) + 2 * lane;
      if (bp + 1 < cap) { out[bp] = (u8)(fin >> 8); out[bp + 1] = (u8)fin; } else overflow = true;
    }
    const u32 moved = __shfl_sync(FULL, fin, (lane + nf) & 31);
    __syncwarp();
    cells[lane] = (lane + nf < 32) ? moved : 0u;
    base += 16 * nf;
    __syncwarp();
    mine = nxt;
  }
  // flush_and_get_buffer (arithmetic.rs:176-195): the first pos + 4 bytes of the number
  const u32 pos = T < 24 ? 0u : 1u + (T - 24) / 8;
  const u32 total_bytes = pos + 4;
  {
    const u32 fin = cells[lane];  // already normalised
    const u32 bp = (base >> 3) + 2 * lane;
    if (bp < total_bytes) { if (bp < cap) out[bp] = (u8)(fin >> 8); else overflow = true; }
    if (bp + 1 < total_bytes) { if (bp + 1 < cap) out[bp + 1] = (u8)fin; else overflow = true; }
  }
  overflow = __any_sync(FULL, overflow);
  if (lane == 0) {
    if (is_hdr) IS.part0_bytes = total_bytes; else IS.part1_bytes = total_bytes;
    if (overflow) IS.status = 4;  // ZW_ERR_OUTPUT_TOO_SMALL (cannot happen with the 7-bits-per-symbol bound)
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
