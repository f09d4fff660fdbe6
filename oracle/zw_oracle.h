/* zw_oracle.h -- C interface of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY. This is a plain C++ restatement of the reference's lossy VP8
 * encode path (imazen/image-webp a.k.a. zenwebp 0.2.0, Rust) used as the parity checker for
 * the CUDA library.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Nothing under image_webp_b200/ includes, links or calls it.
 *
 * PARITY STATUS: "parity unpinned" at file level -- the reference has no golden .webp / hash
 * for its encoder output and no Rust toolchain exists in the build image, so whole-file bytes
 * cannot be compared with the real reference.  Component arithmetic IS pinned by the
 * reference's own known-answer tests (bool coder bytes, trellis-vs-libwebp vector, DCT/IDCT
 * round trip, predictor KATs, fixed-cost constants, lambda formulas): see tests/test_oracle_kats.py.
 */
#ifndef ZW_ORACLE_H
#define ZW_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-macroblock parity record (P1MB / P2MB). 832 bytes.
 * The CUDA library defines the same POD layout independently (zw_device.cuh). */
typedef struct zwo_mb_record {
  uint8_t ymode;       /* 0 DC, 1 V, 2 H, 3 TM, 4 B_PRED                       */
  uint8_t uvmode;      /* 0 DC, 1 V, 2 H, 3 TM                                 */
  uint8_t segment;     /* segment id (0 when segments are disabled)            */
  uint8_t skip;        /* all simple-quantised levels zero                     */
  uint8_t bmodes[16];  /* sub-block modes when ymode==4 (else 0)               */
  uint16_t top_nz;     /* incoming top complexity: bit0 y2, 1..4 y, 5..6 u, 7..8 v */
  uint16_t left_nz;    /* incoming left complexity, same packing               */
  int8_t derr_left[4]; /* left_derr after this MB  [u0,u1,v0,v1]               */
  int8_t derr_top[4];  /* top_derr[mbx] after this MB [u0,u1,v0,v1]            */
  int16_t levels[25][16]; /* coded levels, zig-zag order: [0] Y2, [1..16] Y, [17..20] U, [21..24] V */
} zwo_mb_record;

typedef struct zwo_dump zwo_dump; /* opaque bag of named stage dumps */

/* Encode one image to a raw VP8 key-frame payload (what encode_frame_lossy appends,
 * reference src/encoder/vp8.rs:3132-3153).  color: 0 L8, 1 La8, 2 Rgb8, 3 Rgba8.
 * Returns 0 OK, 1 InvalidDimensions, 2 InvalidBufferSize (reference panics), 3 bad quality.
 * *out is malloc'ed; release with zwo_free.  dump may be NULL. */
int zwo_encode_vp8(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height,
                   int color, int quality, int method, uint8_t** out, size_t* out_len,
                   zwo_dump* dump);

/* Same plus the simple RIFF/WEBP/"VP8 " container (reference src/encoder/api.rs:1320-1329). */
int zwo_encode_webp(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height,
                    int color, int quality, int method, uint8_t** out, size_t* out_len,
                    zwo_dump* dump);
void zwo_free(void* p);

/* VP8 key-frame decoder restated from src/decoder/ (zw_dec_oracle.inc): checker of the on-device decoder. */
int zwo_decode(const uint8_t* data, size_t len, int fancy, uint8_t** rgb, uint8_t** planes, uint8_t** planes_unfiltered,
               uint8_t** mbinfo, uint32_t hdr[16]);

/* Encode n same-sized images with `threads` host threads (one image per thread at a time);
 * returns total output bytes. Used by bench.py's CPU baseline. outs/out_lens may be NULL. */
size_t zwo_encode_batch_mt(const uint8_t* data, size_t n, uint32_t width, uint32_t height,
                           int quality, int method, int threads);

/* Same for any colour type, keeping the outputs: file i is copied to out + i*out_stride (truncated there if
 * longer), out_lens[i] = its full length (0 on failure).  container != 0 adds the RIFF wrap. */
size_t zwo_encode_batch_mt_out(const uint8_t* data, size_t n, uint32_t width, uint32_t height, int color,
                               int quality, int method, int threads, int container, uint8_t* out,
                               size_t out_stride, uint32_t* out_lens);

zwo_dump* zwo_dump_new(void);
void zwo_dump_free(zwo_dump* d);
/* Returns 1 and sets ptr/len (bytes) if a stage called `name` was recorded, else 0. */
int zwo_dump_get(const zwo_dump* d, const char* name, const uint8_t** ptr, size_t* len);

/* ---- component entry points for the reference's known-answer tests ---- */
void zwo_dct4x4(int32_t block[16]);  /* src/common/transform.rs:176 */
void zwo_idct4x4(int32_t block[16]); /* src/common/transform.rs:35  */
void zwo_wht4x4(int32_t block[16]);  /* src/common/transform.rs:116 */
void zwo_iwht4x4(int32_t block[16]); /* src/common/transform.rs:82  */
/* Bool coder session: ops is a sequence of (bit, prob) pairs; out must hold n/8+8 bytes. */
size_t zwo_bool_encode(const uint8_t* bits, const uint8_t* probs, size_t n, uint8_t* out);
/* write_with_tree_start_index on one of the reference's trees:
 * tree_id 0 DCT token, 1 kf ymode, 2 kf bmode, 3 uv mode, 4 segment id. */
size_t zwo_bool_encode_tree(int tree_id, const uint8_t* probs, const int8_t* values, size_t n,
                            int start_index, uint8_t* out);
/* Trellis with an explicitly given matrix (q, iq, bias, sharpen as in VP8Matrix) and
 * level-cost tables computed from the given 1056 token probabilities. */
int zwo_trellis(int32_t coeffs[16], int32_t out[16], const uint16_t q[16], const uint32_t iq[16],
                const uint32_t bias[16], const uint16_t sharpen[16], uint32_t lambda, int first,
                const uint8_t* probs1056, int ctype, int ctx0);
/* Build a VP8Matrix like VP8Matrix::new (cost.rs:401). type: 0 Y1, 1 Y2, 2 UV. */
void zwo_matrix_new(int q_dc, int q_ac, int type, uint16_t q[16], uint32_t iq[16],
                    uint32_t bias[16], uint32_t zthresh[16], uint16_t sharpen[16]);
/* Quality -> quant index (vp8.rs:37-55) and segment params for an index (types.rs:806-853).
 * lambdas[8] = i4, i16, uv, mode, trellis_i4, trellis_i16, trellis_uv, tlambda. */
int zwo_quality_to_quant_index(int quality);
void zwo_segment_lambdas(int quant_index, uint32_t lambdas[8], int16_t quants[6]);
int zwo_compute_segment_quant(int base_quant, int segment_alpha, int sns_strength);
int zwo_compute_filter_level(int quant_index, int sharpness, int filter_strength);
double zwo_cbrt(double x);
double zwo_pow(double x, double n);
/* 4x4 intra predictors on a bordered work buffer, stride 32 (prediction.rs:326-554):
 * mode 0 DC,1 TM,2 VE,3 HE,4 LD,5 RD,6 VR,7 VL,8 HD,9 HU. */
void zwo_predict4x4(uint8_t* ws, int mode, int x0, int y0, int stride);
void zwo_predict4x4_all(const uint8_t* ws, int x0, int y0, int stride, uint8_t out[160]);
void zwo_add_residue(uint8_t* pblock, const int32_t rblock[16], int y0, int x0, int stride);
uint32_t zwo_rd_score(uint32_t sse, uint16_t mode_cost, uint32_t lambda, uint64_t* full);
/* record_coeffs into a fresh 1056-entry stats array (cost.rs:1297). */
void zwo_record_coeffs(const int32_t coeffs[16], int token_type, int first, int ctx,
                       uint32_t stats[1056]);
/* Level-cost tables from probabilities: level_cost[4][8][3][68] (cost.rs:1500). */
void zwo_level_costs(const uint8_t* probs1056, uint16_t* level_cost_6528);
uint32_t zwo_residual_cost(const int32_t levels[16], int ctype, int first, int ctx0,
                           const uint8_t* probs1056, int zero_tables);
int zwo_tdisto_16x16(const uint8_t* a, const uint8_t* b, int stride);
/* RGB->YUV420 only (decoder/yuv.rs:656): planes must hold 16mbw*16mbh and 2x 8mbw*8mbh. */
void zwo_convert_yuv(const uint8_t* rgb, uint32_t w, uint32_t h, int bpp, uint8_t* y, uint8_t* u,
                     uint8_t* v);
/* Fixed table accessors (tests check the spot values cost.rs:2036-2068 pins). */
uint16_t zwo_fixed_cost_i16(int i);
uint16_t zwo_fixed_cost_uv(int i);
uint16_t zwo_fixed_cost_i4(int top, int left, int mode);
uint16_t zwo_entropy_cost(int p);
uint16_t zwo_level_fixed_cost(int level);

/* ---- VP8L lossless encoder and the complete WebPEncoder::encode container logic (zw_lossless_oracle.inc) ----
 * encode_frame_lossless (src/encoder/api.rs:945-1167): the raw VP8L stream.  color as above; returns 0 OK,
 * 1 InvalidDimensions (w or h == 0 or > 16384), 2 InvalidBufferSize (assert_eq! panic in the reference). */
int zwo_encode_lossless(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height, int color, int use_predictor,
                        int implicit_dimensions, uint8_t** out, size_t* out_len);
/* encode_alpha_lossless (api.rs:1175-1222): the payload of an ALPH chunk (La8 / Rgba8 input). */
int zwo_encode_alpha_lossless(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height, int color, uint8_t** out,
                              size_t* out_len);
/* WebPEncoder::encode (api.rs:1291-1394): lossy or lossless frame, simple or VP8X container, ICCP / ALPH / EXIF / XMP. */
int zwo_webp_encode(const uint8_t* data, size_t data_len, uint32_t width, uint32_t height, int color, int use_predictor,
                    int use_lossy, int quality, int method, const uint8_t* icc, size_t icc_len, const uint8_t* exif,
                    size_t exif_len, const uint8_t* xmp, size_t xmp_len, uint8_t** out, size_t* out_len);
size_t zwo_webp_encode_batch_mt(const uint8_t* data, size_t n, uint32_t width, uint32_t height, int color, int use_predictor,
                                int use_lossy, int quality, int method, int threads, uint8_t* out, size_t out_stride,
                                uint32_t* out_lens);
/* build_huffman_tree (api.rs:163-287); returns 0 when at most one symbol is used. */
int zwo_build_huffman(const uint32_t* frequencies, size_t n, int length_limit, uint8_t* lengths, uint16_t* codes);

/* Primitive-invocation counters of the calling thread (measurement: SURVEY.md 8(d) algorithmic
   int-ops).  Order: fdct, idct, wht, iwht, ttransform, quantised coefficients, sse pixels,
   residual-cost coefficients visited, trellis positions, I4 predictor sets, add_residue blocks,
   trellis blocks.  Counted inside the mode-search / transform functions of both passes only.
   zwo_opcounts_get fills out[set * n + i], set = pass-1 luma, pass-1 chroma, pass-2 luma, pass-2
   chroma, and returns n (the number of counters per set). */
void zwo_opcounts_reset(void);
size_t zwo_opcounts_get(uint64_t* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* ZW_ORACLE_H */
