/* zenwebp_b200.h -- C ABI of the B200-native WebP encoder core (lossy VP8 key frames; lossless VP8L; containers).
 *
 * This is the drop-in boundary for ONE path of imazen/image-webp (crate `zenwebp` 0.2.0):
 *     WebPEncoder::encode  ->  encode_frame_lossy(&mut Vec<u8>, &[u8], w, h, ColorType, quality, method)
 *     reference: src/encoder/api.rs:1291-1329 (caller, RIFF wrap), src/encoder/vp8.rs:3132-3153 (seam)
 * The reference has no FFI of its own (it is #![forbid(unsafe_code)] Rust); these entry points are
 * what a `zenwebp-b200-sys` crate binds (see INTEGRATION.md for the Rust / ctypes stubs).
 * Next to it: an on-device VP8 key-frame DECODER used as the batch verifier (zw_decode_batch / zw_verify below).
 *
 * Everything device-side is hand-written CUDA for sm_100a.  There is NO CPU fallback: every
 * call fails with ZW_ERR_CUDA (>= 100) when no CUDA device / driver is usable.
 * Plain C types only; no torch, no C++ in the signatures.
 */
#ifndef ZENWEBP_B200_H
#define ZENWEBP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes.  0..3 mirror the reference's EncodingError / panics
 * (src/encoder/api.rs:35-48, src/encoder/vp8.rs:1307-1313, :2401-2403, :3143-3148). */
enum {
  ZW_OK = 0,
  ZW_ERR_INVALID_DIMENSIONS = 1, /* EncodingError::InvalidDimensions (w or h == 0 or > 16383)            */
  ZW_ERR_INVALID_BUFFER_SIZE = 2,/* w*h*bpp != len  (reference: assert_eq! panic / InvalidBufferSize)     */
  ZW_ERR_INVALID_PARAM = 3,      /* quality > 100 (reference panics); null pointers; unsupported colour   */
  ZW_ERR_OUTPUT_TOO_SMALL = 4,   /* caller-provided output buffer smaller than the bitstream               */
  ZW_ERR_PARTITION_TOO_LARGE = 5,/* first partition >= 2^19 bytes: the 19-bit size field would overflow
                                    (the reference silently emits a corrupt tag here, vp8.rs:320)         */
  ZW_ERR_NOT_STAGED = 6,         /* zw_encode_resident / zw_download without a staged batch; unknown ticket */
  ZW_ERR_BUSY = 7,               /* zw_submit: every pipeline slot holds a batch (zw_wait + zw_release one)  */
  ZW_ERR_TOO_LARGE = 8,          /* zw_submit / zw_stage_batch: batch exceeds max_device_bytes or 65535
                                    images (zw_encode_*_batch split such batches into chunks themselves)  */
  ZW_ERR_CUDA = 100              /* 100 + cudaError_t                                                      */
};

/* ColorType (src/encoder/api.rs:83-92).  All four are accepted: L8/La8 go through convert_image_y
 * (decoder/yuv.rs:806: Y = the grey sample, U = V = 127); the alpha of La8/Rgba8 is ignored by the
 * VP8 payload exactly as vp8.rs:1296 does (see zw_encode_webp_batch for the container). */
enum { ZW_COLOR_L8 = 0, ZW_COLOR_LA8 = 1, ZW_COLOR_RGB8 = 2, ZW_COLOR_RGBA8 = 3 };

typedef struct zw_ctx zw_ctx;

typedef struct zw_limits {
  size_t max_device_bytes; /* working-set budget per chunk of a batch; 0 = default (32 GiB)     */
  int persistent_warps_per_sm; /* 0 = default; tuning knob of the wavefront kernels            */
  int reserved[5];         /* reserved[0]: pipeline depth = batches (chunks) one context keeps in flight,
                              each on its own stream with its own buffers, so that the H2D / D2H
                              copies of one hide behind the kernels of another; 0 = default (3),
                              max 8.  Buffers are allocated on first use.  Others must be 0.     */
} zw_limits;

/* One input image: caller-owned, tightly packed rows (stride = width * bpp), host memory
 * (pinned memory makes the H2D copy asynchronous; pageable memory also works).  Images that are
 * contiguous in host memory (image i+1 starts where image i ends, lengths multiples of 16) are
 * copied with one transfer per run.
 * LIFETIME: the buffers must stay valid and unchanged until the call that received them returns
 * (zw_encode_*_batch, zw_multi_encode, zw_stage_batch) or, for zw_submit, until zw_wait on its ticket
 * has returned. */
typedef struct zw_image {
  const uint8_t* data;
  size_t len;      /* bytes in data; must equal width*height*bpp                                  */
  uint32_t width;  /* 1..16383                                                                     */
  uint32_t height; /* 1..16383                                                                     */
  uint32_t color;  /* ZW_COLOR_*                                                                   */
  uint32_t reserved;
} zw_image;

/* One output slot.  If data == NULL the library allocates (malloc) and the caller releases with
 * zw_free; otherwise `cap` bytes at `data` are caller-owned and `len` receives the size. */
typedef struct zw_output {
  uint8_t* data;
  size_t cap;
  size_t len;
  int status; /* per-image ZW_* code */
  int reserved;
} zw_output;

/* Per-stage device times (CUDA events on the library's stream), milliseconds, summed over chunks. */
typedef struct zw_timing {
  float h2d_ms, yuv_ms, analysis_ms, pass1_ms, stats_ms, pass2_ms, token_ms, boolcode_ms, assemble_ms, d2h_ms;
  float device_total_ms; /* first kernel .. last kernel                                           */
  float wall_ms;         /* host wall clock of the whole call                                     */
  uint64_t kernel_launches;
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t pixels;
  float chroma1_ms, chroma2_ms; /* pass-1 chroma chain / pass-2 chroma wavefront (pass1_ms and
                                   pass2_ms time the luma wavefront kernels alone)             */
  uint64_t symbols;             /* boolean-coder input symbols of both partitions, all images    */
} zw_timing;

/* Create / destroy an encoder context bound to one CUDA device.  One context per (host thread,
 * GPU); a context is not thread-safe, contexts are independent.  Returns NULL on failure
 * (zw_last_error() tells why). */
zw_ctx* zw_create(int device, const zw_limits* limits);
void zw_destroy(zw_ctx* ctx);
int zw_last_error(void);
const char* zw_strerror(int code);
void zw_free(void* p);
/* Conservative bound for one output (payload + RIFF header). */
size_t zw_max_output_size(uint32_t width, uint32_t height);

/* Batch entry: n independent images -> n raw VP8 key-frame payloads, byte-identical to the CPU port
 * (oracle/) of what the reference's encode_frame_lossy appends (vp8.rs:3132).  quality 0..100,
 * method 0..6 (larger values clamped like vp8.rs:1291).  The batch is split into chunks (device
 * budget; ~4 chunks for large batches) that are pipelined over the context's slots, so the copies
 * of one chunk hide behind the kernels of another.  Returns ZW_OK if the call ran; per-image results
 * are in outs[i].status (every outs[i] is initialised even when the call fails). */
int zw_encode_vp8_batch(zw_ctx* ctx, const zw_image* imgs, size_t n, int quality, int method,
                        zw_output* outs, zw_timing* timing);

/* Same, wrapped in the simple RIFF container exactly as WebPEncoder::encode does for opaque
 * input without metadata (api.rs:1320-1329): these bytes ARE the .webp file.  This entry (like zw_submit /
 * zw_wait and zw_multi_encode) writes the simple container only: for the alpha colour types (ZW_COLOR_LA8 /
 * ZW_COLOR_RGBA8) the reference writes VP8X + a lossless ALPH chunk instead (api.rs:1330-1394), so images of those
 * types get outs[i].status = ZW_ERR_INVALID_PARAM here -- zw_encode_batch below builds that file. */
int zw_encode_webp_batch(zw_ctx* ctx, const zw_image* imgs, size_t n, int quality, int method,
                         zw_output* outs, zw_timing* timing);

/* Streaming form: one context keeps up to `depth` batches in flight (zw_limits.reserved[0]).
 *   zw_submit   validate + start the H2D copy + enqueue every kernel of one batch; returns at once with
 *               a ticket.  No host synchronisation happens inside a batch: stream offsets are computed on
 *               the device (k_layout), the symbol arenas are sized by estimate and checked on the device.
 *               ZW_ERR_BUSY when every slot is taken, ZW_ERR_TOO_LARGE when the batch exceeds the budget.
 *   zw_wait     block until that batch is finished, fetch all files with ONE D2H copy into a pinned arena
 *               owned by the slot, and describe them in *view (no per-image allocation, no copy):
 *               file i = view->arena[view->offsets[i] .. + view->lens[i]), view->status[i] = ZW_* code.
 *               container != 0: the .webp file (RIFF wrap as api.rs:1325-1329); 0: the raw VP8 payload.
 *               May be called again (e.g. with the other container flag) until the ticket is released.
 *   zw_release  give the slot back; the view of that ticket becomes invalid.
 * Batches submitted to one context run their kernels in submission order; the copies of a batch overlap
 * the kernels of its neighbours.  Bytes are identical to zw_encode_*_batch. */
typedef struct zw_batch_view {
  const uint8_t* arena;    /* pinned host memory owned by the slot, valid until zw_release(ticket)   */
  size_t n;                /* images of the batch                                                       */
  const uint64_t* offsets; /* [n] byte offset of file i in arena                                        */
  const uint32_t* lens;    /* [n] bytes of file i (0 when status[i] != ZW_OK)                           */
  const int32_t* status;   /* [n] per-image ZW_* code                                                   */
} zw_batch_view;
int zw_submit(zw_ctx* ctx, const zw_image* imgs, size_t n, int quality, int method, int* ticket);
int zw_wait(zw_ctx* ctx, int ticket, int container, zw_batch_view* view, zw_timing* timing);
int zw_release(zw_ctx* ctx, int ticket);

/* One batch over several GPUs of one box (north_star "sharded by image across the 8
 * GPUs ... gathered to the host"): contiguous slices of ceil(n / G) images, one host thread + context per
 * GPU (created once by zw_multi_create), no collective on the data path; outs[] is filled in image order.
 * per_device (NULL or [device count]) receives each GPU's timing.  container as in zw_wait. */
typedef struct zw_multi zw_multi;
zw_multi* zw_multi_create(const int* devices, int n_devices, const zw_limits* limits);
void zw_multi_destroy(zw_multi* m);
int zw_multi_device_count(const zw_multi* m);
int zw_multi_encode(zw_multi* m, const zw_image* imgs, size_t n, int quality, int method, int container,
                    zw_output* outs, zw_timing* per_device);

/* Split form, for callers that keep inputs resident in HBM (and for kernel-only timing):
 *   zw_stage_batch     validate + H2D copy of one chunk (must fit max_device_bytes); returns when
 *                      the copy has finished (the caller may reuse its buffers)
 *   zw_encode_resident run every kernel on the staged inputs; bitstreams stay on the device
 *   zw_download        D2H + host chunk assembly into outs (container != 0 adds the RIFF wrap) */
int zw_stage_batch(zw_ctx* ctx, const zw_image* imgs, size_t n);
int zw_encode_resident(zw_ctx* ctx, int quality, int method, zw_timing* timing);
int zw_download(zw_ctx* ctx, zw_output* outs, size_t n, int container, zw_timing* timing);

/* Measurement helper (no reference counterpart; SURVEY.md 8(d) "INT peak: measure, don't assume"):
 * runs an integer-issue microbenchmark (IMAD + LOP3/IADD3 chains, full occupancy) on the context's
 * GPU and returns thread-level integer instructions per second -- the denominator of the
 * mode-search kernels' integer roofline in bench.py. */
int zw_measure_int_peak(zw_ctx* ctx, double* int_instr_per_s);

/* Parity/debug: copy a named intermediate stage of image `index` of the last encoded chunk to
 * host memory (names follow SURVEY.md Appendix F: "YUV_Y","YUV_U","YUV_V","ALPHA","ALPHA_HIST",
 * "SEG_MAP","SEG_QIDX","SEG_TREE_PROBS","SEG_UPDATE_MAP","P1MB","STATS","PROBS","SKIP_PROB",
 * "LCOST","P2MB","PART0","PART1","VP8","WEBP").  For a batch that was split into chunks, `index`
 * counts inside the LAST chunk.  *len receives the stage size; if cap is too small
 * nothing is copied and ZW_ERR_OUTPUT_TOO_SMALL is returned. */
int zw_dump_stage(zw_ctx* ctx, size_t index, const char* stage, void* dst, size_t cap, size_t* len);

/* ---- On-device VP8 key-frame decoder, used as the batch verifier --------------------------
 * Reference: Vp8Decoder::decode_frame (src/decoder/vp8.rs:1526) + the loop filter (loop_filter.rs) + the
 * YUV->RGB conversion WebPDecoder::read_image applies (src/decoder/yuv.rs:82 bilinear "fancy" upsampling, the
 * default, and :402 UpsamplingMethod::Simple).  Decoded pixels are identical to the CPU port of that decoder
 * (oracle/zw_dec_oracle.inc), which is pinned pixel-exact by the reference's own decode fixtures.
 * Key frames only (every WebP still image); any number of token partitions, both loop filters, segments. */
enum {
  ZW_DEC_OK = 0,
  ZW_DEC_BITSTREAM = 1,   /* DecodingError::BitStreamError: a partition ended early                          */
  ZW_DEC_UNSUPPORTED = 2, /* not a key frame (DecodingError::UnsupportedFeature)                             */
  ZW_DEC_MAGIC = 3,       /* start code != 9d 01 2a (DecodingError::Vp8MagicInvalid)                          */
  ZW_DEC_COLORSPACE = 4,  /* DecodingError::ColorSpaceInvalid                                                */
  ZW_DEC_TRUNCATED = 5,   /* file shorter than its headers / first partition say                            */
  ZW_DEC_CONTAINER = 6,   /* RIFF/WEBP without a 'VP8 ' chunk (lossless VP8L files are not this path)        */
  ZW_DEC_DIMENSIONS = 7   /* zero-sized frame, or the source handed in for verification has another size    */
};
typedef struct zw_blob {
  const uint8_t* data; /* a .webp file (RIFF container, simple or VP8X) or a bare VP8 frame; host memory */
  size_t len;
} zw_blob;
typedef struct zw_decode_info {
  int32_t status; /* ZW_DEC_* */
  uint32_t width, height;
  uint32_t filter_type, filter_level, sharpness, num_partitions, segments_enabled; /* frame header fields */
  uint64_t sse_rgb; /* verification: sum over width*height*3 samples of (decoded - source)^2       */
  double psnr_rgb;  /* verification: 10 log10(255^2 * 3wh / sse_rgb); 99.0 for identical pixels   */
} zw_decode_info;

/* Decode n files on the GPU: the serial bitstream walk (one warp per image), then reconstruction and loop filter
 * as wavefronts over macroblock rows, then a data-parallel colour conversion.  upsampling: 1 = bilinear (the reference's default), 0 = nearest.
 *   rgb_outs  NULL, or n slots that receive width*height*3 RGB bytes (allocated with malloc when data == NULL)
 *   sources   NULL, or n source images: every decoded image is scored against sources[i] on the device
 *             (sse_rgb / psnr_rgb; grey sources compare each channel with the grey value, alpha is ignored)
 *   infos     n results;  device_ms NULL or [4]: parse, reconstruct, filter, colour kernels (CUDA events) */
int zw_decode_batch(zw_ctx* ctx, const zw_blob* files, size_t n, int upsampling, zw_output* rgb_outs,
                    const zw_image* sources, zw_decode_info* infos, float* device_ms);

/* Verify a batch where it lies: decode the files zw_submit left in device memory and score them against the
 * batch's source pixels, which are still resident there as well -- only the per-image results cross the link.
 * Call between zw_submit (any time: it waits for the batch's kernels) and zw_release of that ticket; infos has
 * one entry per image of the batch. */
int zw_verify(zw_ctx* ctx, int ticket, int upsampling, zw_decode_info* infos, float* device_ms);

/* Parity/debug: a stage of image `index` of the last decoded chunk: "DEC_PLANES" (filtered Y | U | V of the
 * padded frame), "DEC_MBINFO" (4 words per macroblock: flags, sub-block modes), "DEC_STATE". */
int zw_decode_dump_stage(zw_ctx* ctx, size_t index, const char* stage, void* dst, size_t cap, size_t* len);

/* ---- Lossless (VP8L) encoder and the complete WebPEncoder::encode (SURVEY.md 8(f)1 + 8(f)4) --------------------------
 * Reference: encode_frame_lossless (src/encoder/api.rs:945-1167: subtract-green and "top" predictor transforms, one
 * Huffman code per channel, run-length back references), encode_alpha_lossless (:1175-1222) and the container logic of
 * WebPEncoder::encode (:1291-1394).  Bytes are identical to the CPU port (oracle/zw_lossless_oracle.inc), whose files
 * libwebp decodes back to exactly the input pixels (the reference's own acceptance test, api.rs:1447-1511).
 * Dimensions 1..16384 (api.rs:968); status codes as above.  The lossless kernels have their own buffers and streams:
 * batches in flight from zw_submit keep running and their tickets stay valid.  (zw_encode_batch with use_lossy goes
 * through the blocking lossy batch call, which -- like zw_encode_vp8_batch / zw_encode_webp_batch -- takes every pipeline
 * slot of the context: wait for and release outstanding tickets first, or they are dropped.)  zw_timing fields used: h2d_ms, yuv_ms (transforms + run heads), analysis_ms
 * (tokens + histograms), stats_ms (Huffman codes), token_ms (bit counts + scan), assemble_ms (bit packing), d2h_ms. */

/* n raw VP8L streams (container == 0) or .webp files in the simple container (container != 0). */
int zw_encode_lossless_batch(zw_ctx* ctx, const zw_image* imgs, size_t n, int use_predictor_transform, int container,
                             zw_output* outs, zw_timing* timing);
/* n ALPH chunk payloads: the alpha channel of ZW_COLOR_LA8 / ZW_COLOR_RGBA8 images (encode_alpha_lossless). */
int zw_encode_alpha_batch(zw_ctx* ctx, const zw_image* imgs, size_t n, zw_output* outs, zw_timing* timing);

/* EncoderParams (api.rs:425-443); zw_params_default() = EncoderParams::default(): lossless, predictor on, q95, m4. */
typedef struct zw_params {
  int use_predictor_transform;
  int use_lossy;
  int lossy_quality; /* 0..100 */
  int method;        /* 0..6 (larger values are clamped) */
} zw_params;
/* WebPEncoder::set_icc_profile / set_exif_metadata / set_xmp_metadata (api.rs:1267-1279); empty = absent. */
typedef struct zw_metadata {
  const uint8_t* icc_profile; size_t icc_len;
  const uint8_t* exif;        size_t exif_len;
  const uint8_t* xmp;         size_t xmp_len;
} zw_metadata;
zw_params zw_params_default(void);
/* WebPEncoder::encode for a batch: lossy ("VP8 ") or lossless ("VP8L") frame; simple container, or the extended one
 * (VP8X + ICCP + ALPH + frame + EXIF + XMP) when an image has metadata or is lossy with an alpha colour type.
 * meta: NULL or n entries.  outs[i] receives the complete .webp file. */
int zw_encode_batch(zw_ctx* ctx, const zw_image* imgs, size_t n, const zw_params* params, const zw_metadata* meta,
                    zw_output* outs, zw_timing* timing);
/* Parity/debug: a stage of image `index` of the last lossless chunk: "LL_RESIDUAL" (u32 per pixel: R-G, G, B-G, A after
 * the transforms), "LL_TOKENS" (u16 per pixel: bit 0 literal, bits 1.. run length carried), "LL_HIST" ([4][280] u32: red,
 * green + lengths, blue, alpha), "LL_CODES" ([4][280] u32: length << 16 | code), "LL_HEADER", "LL_STATE". */
int zw_lossless_dump_stage(zw_ctx* ctx, size_t index, const char* stage, void* dst, size_t cap, size_t* len);

/* Library build info, e.g. "zenwebp_b200 0.2 (CUDA, sm_100a)". */
const char* zw_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ZENWEBP_B200_H */
