// zw_boolcoder.cuh -- the lane-local pieces of the segment-parallel boolean coder (k_bc_* in zw_back.cuh).
//
// ZW_HD (__host__ __device__) only so that tests/hostcheck can run the whole five-step scheme on the CPU against
// the oracle's ArithmeticEncoder; the shipped library only calls these from kernels.
// Reference: src/encoder/arithmetic.rs:7-196 (write_bool :67-95, add_one_to_output :47-60,
// flush_and_get_buffer :176-195).
#ifndef ZW_BOOLCODER_CUH
#define ZW_BOOLCODER_CUH
#include "zw_prims.cuh"

namespace zw {

// A symbol stream is cut into segments of BC_SEG symbols that are coded in parallel.
#ifndef ZW_BC_SEG
#define ZW_BC_SEG 8192
#endif
#ifndef ZW_BC_WARM
#define ZW_BC_WARM 1024
#endif
constexpr u32 BC_SEG = ZW_BC_SEG;    // symbols per segment (multiple of 8: segments start 16-byte aligned)
constexpr u32 BC_WARM = ZW_BC_WARM;  // symbols before a segment start over which the possible range states are narrowed down
ZW_HD u32 bc_segments(u32 n_symbols) { return n_symbols == 0 ? 1u : (n_symbols + BC_SEG - 1) / BC_SEG; }

ZW_HD u32 bc_clz32(u32 v) {
#if defined(__CUDA_ARCH__)
  return (u32)__clz((int)v);
#else
  return (u32)__builtin_clz(v);
#endif
}
ZW_HD u32 bc_popc(u32 v) {
#if defined(__CUDA_ARCH__)
  return (u32)__popc(v);
#else
  return (u32)__builtin_popcount(v);
#endif
}
ZW_HD u32 bc_fns(u32 m, u32 r) {  // position of the r-th set bit (r from 0) of m
#if defined(__CUDA_ARCH__)
  return __fns(m, 0, (int)r + 1);
#else
  for (u32 b = 0; b < 32; b++)
    if ((m >> b) & 1u) { if (r == 0) return b; r--; }
  return 0xffffffffu;
#endif
}

// One step of the range recurrence on rm1 = range - 1 (in [127, 254]); sym = bit << 8 | prob.  Returns the new
// rm1; `add` = what write_bool adds to bottom (split for a 1, else 0), `shift` = the renormalisation shifts.
ZW_HD u32 bc_step(u32 rm1, u32 sym, u32& add, u32& shift) {
  const u32 x = (rm1 * (sym & 255u)) >> 8;  // split - 1
  const bool bit = (sym >> 8) != 0;
  const u32 r2 = bit ? rm1 - x : x + 1;     // new range, 1..255
  add = bit ? x + 1 : 0u;
  shift = bc_clz32(r2) - 24u;
  return (r2 << shift) - 1u;
}

ZW_HD u32 bc_sym8(u32 x, u32 y, u32 z, u32 w, int e) {  // symbol e (0..7) of eight packed u16 (one 16-byte load)
  const u32 v = e < 2 ? x : (e < 4 ? y : (e < 6 ? z : w));
  return (e & 1) ? (v >> 16) : (v & 0xffffu);
}

// position (0..127) of the r-th set bit (r from 0) of the 128-bit map m0..m3 (scalars: no local-memory indexing)
ZW_HD u32 bc_nth_bit(u32 m0, u32 m1, u32 m2, u32 m3, u32 r) {
  const u32 c0 = bc_popc(m0), c1 = bc_popc(m1), c2 = bc_popc(m2);
  u32 w, m;
  if (r < c0) { w = 0; m = m0; }
  else if (r < c0 + c1) { w = 1; m = m1; r -= c0; }
  else if (r < c0 + c1 + c2) { w = 2; m = m2; r -= c0 + c1; }
  else { w = 3; m = m3; r -= c0 + c1 + c2; }
  return w * 32u + bc_fns(m, r);
}
ZW_HD u32 bc_rank(u32 m0, u32 m1, u32 m2, u32 m3, u32 bit) {  // set bits below `bit`
  const u32 w = bit >> 5, below = (1u << (bit & 31u)) - 1u;
  u32 r = bc_popc((w == 0 ? m0 : (w == 1 ? m1 : (w == 2 ? m2 : m3))) & below);
  if (w > 0) r += bc_popc(m0);
  if (w > 1) r += bc_popc(m1);
  if (w > 2) r += bc_popc(m2);
  return r;
}

// add_one_to_output (arithmetic.rs:47-60): ripple a carry through trailing 0xFF bytes below `pos`, not below `floor`.
// Returns false when the carry leaves [floor, pos).
ZW_HD bool bool_carry(u8* out, u32 pos, u32 floor) {
  u32 j = pos;
  while (j > floor) {
    j--;
    if (out[j] < 255) { out[j]++; return true; }
    out[j] = 0;
  }
  return false;
}

// bytes written and bits to go until the next byte completes, after `bits` renormalisation shifts
// (the coder starts with bit_num = 24 and emits its first byte after 24 shifts)
ZW_HD void bc_position(u64 bits, u32& pos, int& bit_num) {
  if (bits < 24) { pos = 0; bit_num = 24 - (int)bits; }
  else { pos = 1u + (u32)((bits - 24) >> 3); bit_num = 8 - (int)((bits - 24) & 7); }
}

// The reference's serial coder for ONE segment, started in the middle of a stream: write_bool (arithmetic.rs:67-95)
// with the bit-at-a-time renormalisation loop collapsed into at most two steps around the byte boundary, from
// (range = state + 1, bottom = 0, the bit position the segment starts at).  Bytes go straight to their final place
// out[pos]; a carry that would leave the segment's own bytes [own, pos) is counted instead.
struct BcCoder {
  u8* out;
  u32 cap, rm1, bottom, pos, own, carries;
  int bit_num;
  bool overflow;
  ZW_HD void begin(u32 state, u64 start_bit, u8* out_, u32 cap_) {
    out = out_; cap = cap_; rm1 = state; bottom = 0; carries = 0; overflow = false;
    bc_position(start_bit, pos, bit_num);
    own = pos;
  }
  ZW_HD void put(u32 sym) {
    u32 add, sh;
    rm1 = bc_step(rm1, sym, add, sh);
    bottom += add;
    int s2 = (int)sh;
    if (s2 >= bit_num) {  // a byte completes inside this renormalisation (bit_num <= 7 here); bit 32 - bit_num is the carry
      if (bottom >> (32 - bit_num)) { if (!bool_carry(out, pos, own)) carries++; }
      bottom <<= bit_num;
      if (pos < cap) out[pos] = (u8)(bottom >> 24); else overflow = true;
      pos++;
      bottom &= 0xffffffu;
      s2 -= bit_num;
      bit_num = 8;
    }
    bottom <<= s2;
    bit_num -= s2;
  }
  // what is still pending when the segment ends: byte `pos` in bits 24..31, a pending carry in bit 32
  ZW_HD u64 tail() const { return (u64)bottom << bit_num; }
  // flush_and_get_buffer (arithmetic.rs:176-195): carry check, then the four bytes of bottom << bit_num.  Returns the stream length.
  ZW_HD u32 flush() {
    if (bottom & (1u << (32 - bit_num))) { if (!bool_carry(out, pos, own)) carries++; }
    const u32 v = bottom << bit_num;
    for (int k = 0; k < 4; k++) {
      if (pos + k < cap) out[pos + k] = (u8)(v >> (24 - 8 * k)); else overflow = true;
    }
    pos += 4;
    return pos;
  }
};

// Fix-up at the boundary in front of a segment: add the previous segment's tail to the four bytes the segment starts
// with (big endian) and ripple what overflows, plus the carries the segment itself pushed out, into earlier bytes.
ZW_HD void bc_fix_boundary(u8* out, u64 start_bit, u64 prev_tail, u32 carries) {
  u32 pos;
  int bn;
  bc_position(start_bit, pos, bn);
  u8* o = out + pos;
  const u64 sum = (u64)(((u32)o[0] << 24) | ((u32)o[1] << 16) | ((u32)o[2] << 8) | (u32)o[3]) + (prev_tail & 0xffffffffull);
  o[0] = (u8)(sum >> 24); o[1] = (u8)(sum >> 16); o[2] = (u8)(sum >> 8); o[3] = (u8)sum;
  u32 c = (u32)(sum >> 32) + (u32)(prev_tail >> 32) + carries;
  for (; c > 0; c--) bool_carry(out, pos, 0);
}

}  // namespace zw
#endif
