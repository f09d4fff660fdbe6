// zenwebp_b200.hpp -- C++ host-side mirror of the reference's encoder API for the lossy path,
// header-only on top of the C ABI (zenwebp_b200.h).  Mirrors imazen/image-webp `zenwebp` 0.2.0:
//   ColorType      src/encoder/api.rs:83-92      EncodingError  src/encoder/api.rs:35-48
//   EncoderParams  src/encoder/api.rs:419-459    WebPEncoder    src/encoder/api.rs:1244-1398
// plus the batch entry point.  The reference is Rust; with no Rust toolchain in the build image
// this header (and the Python mirror) stand where the wrapper crate of rust/zenwebp-b200 would.
#ifndef ZENWEBP_B200_HPP
#define ZENWEBP_B200_HPP
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zenwebp_b200.h"

namespace zenwebp_b200 {

enum class ColorType { L8 = ZW_COLOR_L8, La8 = ZW_COLOR_LA8, Rgb8 = ZW_COLOR_RGB8, Rgba8 = ZW_COLOR_RGBA8 };

struct EncodingError : std::runtime_error {
  int code;
  EncodingError(int c) : std::runtime_error(zw_strerror(c)), code(c) {}
};
struct InvalidDimensions : EncodingError { InvalidDimensions() : EncodingError(ZW_ERR_INVALID_DIMENSIONS) {} };
struct InvalidBufferSize : EncodingError { InvalidBufferSize() : EncodingError(ZW_ERR_INVALID_BUFFER_SIZE) {} };

struct EncoderParams {
  bool use_predictor_transform = true;
  bool use_lossy = false;
  uint8_t lossy_quality = 95;
  uint8_t method = 4;
  static EncoderParams lossless() { return EncoderParams(); }
  static EncoderParams lossy(uint8_t quality) { EncoderParams p; p.use_lossy = true; p.lossy_quality = quality; return p; }
  EncoderParams with_method(uint8_t m) const { EncoderParams p = *this; p.method = m; return p; }
};

inline void raise_for(int status) {
  if (status == ZW_OK) return;
  if (status == ZW_ERR_INVALID_DIMENSIONS) throw InvalidDimensions();
  if (status == ZW_ERR_INVALID_BUFFER_SIZE) throw InvalidBufferSize();
  throw EncodingError(status);
}

// One context per (host thread, GPU).
class Context {
 public:
  explicit Context(int device = 0) : h_(zw_create(device, nullptr)) { if (!h_) throw EncodingError(zw_last_error()); }
  ~Context() { zw_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  struct ImageRef { const uint8_t* data; size_t len; uint32_t width, height; ColorType color; };
  // Batch entry: n images -> n .webp files (per-image errors throw for the first failing image).
  std::vector<std::vector<uint8_t>> encode_batch(const std::vector<ImageRef>& imgs, const EncoderParams& p, zw_timing* t = nullptr) {
    if (!p.use_lossy) throw std::logic_error("only the lossy VP8 path is implemented on the GPU");
    std::vector<zw_image> in(imgs.size());
    std::vector<zw_output> out(imgs.size());
    for (size_t i = 0; i < imgs.size(); i++) {
      in[i] = zw_image{imgs[i].data, imgs[i].len, imgs[i].width, imgs[i].height, (uint32_t)imgs[i].color, 0};
      out[i] = zw_output{nullptr, 0, 0, 0, 0};
    }
    const int rc = zw_encode_webp_batch(h_, in.data(), in.size(), p.lossy_quality, p.method, out.data(), t);
    std::vector<std::vector<uint8_t>> res(imgs.size());
    int first_err = rc;
    for (size_t i = 0; i < imgs.size(); i++) {
      if (out[i].status == ZW_OK && out[i].data) res[i].assign(out[i].data, out[i].data + out[i].len);
      else if (!first_err) first_err = out[i].status;
      zw_free(out[i].data);
    }
    raise_for(first_err);
    return res;
  }
  // Decoder mirror (WebPDecoder::read_image for lossy still images, src/decoder/api.rs:636): n files -> n RGB images
  // (width * height * 3, bilinear chroma upsampling unless `simple_upsampling`); the first failing file throws.
  struct Decoded { std::vector<uint8_t> rgb; uint32_t width = 0, height = 0; double psnr_rgb = 0.0; };
  std::vector<Decoded> decode_batch(const std::vector<std::pair<const uint8_t*, size_t>>& files, bool simple_upsampling = false,
                                    const std::vector<ImageRef>* sources = nullptr) {
    std::vector<zw_blob> in(files.size());
    std::vector<zw_output> out(files.size(), zw_output{nullptr, 0, 0, 0, 0});
    std::vector<zw_decode_info> info(files.size());
    std::vector<zw_image> src;
    for (size_t i = 0; i < files.size(); i++) in[i] = zw_blob{files[i].first, files[i].second};
    if (sources) for (const ImageRef& r : *sources) src.push_back(zw_image{r.data, r.len, r.width, r.height, (uint32_t)r.color, 0});
    const int rc = zw_decode_batch(h_, in.data(), in.size(), simple_upsampling ? 0 : 1, out.data(), sources ? src.data() : nullptr, info.data(), nullptr);
    std::vector<Decoded> res(files.size());
    int first_err = rc;
    for (size_t i = 0; i < files.size(); i++) {
      if (info[i].status == ZW_DEC_OK && out[i].status == ZW_OK && out[i].data) {
        res[i].rgb.assign(out[i].data, out[i].data + out[i].len);
        res[i].width = info[i].width; res[i].height = info[i].height; res[i].psnr_rgb = info[i].psnr_rgb;
      } else if (!first_err) first_err = info[i].status ? 1000 + info[i].status : out[i].status;
      zw_free(out[i].data);
    }
    if (first_err >= 1000) throw std::runtime_error("zenwebp_b200: decoding error " + std::to_string(first_err - 1000));
    raise_for(first_err);
    return res;
  }
  zw_ctx* handle() { return h_; }
 private:
  zw_ctx* h_;
};

// Streaming batch entry on ONE context (zw_submit / zw_wait / zw_release): up to `depth` batches in flight, the
// H2D copy of a batch running under the kernels of the one before it and its D2H copy under the kernels of the one
// after it.  `submit` returns a ticket object; `get()` blocks until that batch is finished and returns its files.
// The caller must keep the image memory alive until get() returns.  Results equal encode_batch's.
class BatchPipeline {
 public:
  class Pending {
   public:
    Pending(Pending&& o) noexcept : ctx_(o.ctx_), ticket_(o.ticket_) { o.ticket_ = -1; }
    Pending(const Pending&) = delete;
    ~Pending() { if (ticket_ >= 0) zw_release(ctx_, ticket_); }
    std::vector<std::vector<uint8_t>> get(zw_timing* t = nullptr) {
      if (ticket_ < 0) throw std::logic_error("batch already collected");
      zw_batch_view v;
      const int rc = zw_wait(ctx_, ticket_, 1, &v, t);
      if (rc != ZW_OK) { zw_release(ctx_, ticket_); ticket_ = -1; raise_for(rc); }
      std::vector<std::vector<uint8_t>> res(v.n);
      int first_err = 0;
      for (size_t i = 0; i < v.n; i++) {
        if (v.status[i] == ZW_OK) res[i].assign(v.arena + v.offsets[i], v.arena + v.offsets[i] + v.lens[i]);
        else if (!first_err) first_err = v.status[i];
      }
      zw_release(ctx_, ticket_);
      ticket_ = -1;
      raise_for(first_err);
      return res;
    }
   private:
    friend class BatchPipeline;
    Pending(zw_ctx* c, int t) : ctx_(c), ticket_(t) {}
    zw_ctx* ctx_;
    int ticket_;
  };
  explicit BatchPipeline(int device = 0, int depth = 3) {
    zw_limits lim{};
    lim.reserved[0] = depth;
    h_ = zw_create(device, &lim);
    if (!h_) throw EncodingError(zw_last_error());
  }
  ~BatchPipeline() { zw_destroy(h_); }
  BatchPipeline(const BatchPipeline&) = delete;
  BatchPipeline& operator=(const BatchPipeline&) = delete;
  // Throws EncodingError(ZW_ERR_BUSY) when `depth` batches are outstanding: get() one first.
  Pending submit(const std::vector<Context::ImageRef>& imgs, const EncoderParams& p) {
    if (!p.use_lossy) throw std::logic_error("only the lossy VP8 path is implemented on the GPU");
    std::vector<zw_image> in(imgs.size());
    for (size_t i = 0; i < imgs.size(); i++)
      in[i] = zw_image{imgs[i].data, imgs[i].len, imgs[i].width, imgs[i].height, (uint32_t)imgs[i].color, 0};
    int ticket = -1;
    raise_for(zw_submit(h_, in.data(), in.size(), p.lossy_quality, p.method, &ticket));
    return Pending(h_, ticket);
  }
 private:
  zw_ctx* h_;
};

// One batch over several GPUs of one box (zw_multi_*): sharded by image, gathered in image order.
class MultiContext {
 public:
  explicit MultiContext(const std::vector<int>& devices) : h_(zw_multi_create(devices.data(), (int)devices.size(), nullptr)) {
    if (!h_) throw EncodingError(zw_last_error());
  }
  ~MultiContext() { zw_multi_destroy(h_); }
  MultiContext(const MultiContext&) = delete;
  MultiContext& operator=(const MultiContext&) = delete;
  std::vector<std::vector<uint8_t>> encode_batch(const std::vector<Context::ImageRef>& imgs, const EncoderParams& p) {
    if (!p.use_lossy) throw std::logic_error("only the lossy VP8 path is implemented on the GPU");
    std::vector<zw_image> in(imgs.size());
    std::vector<zw_output> out(imgs.size());
    for (size_t i = 0; i < imgs.size(); i++) {
      in[i] = zw_image{imgs[i].data, imgs[i].len, imgs[i].width, imgs[i].height, (uint32_t)imgs[i].color, 0};
      out[i] = zw_output{nullptr, 0, 0, 0, 0};
    }
    const int rc = zw_multi_encode(h_, in.data(), in.size(), p.lossy_quality, p.method, 1, out.data(), nullptr);
    std::vector<std::vector<uint8_t>> res(imgs.size());
    int first_err = rc;
    for (size_t i = 0; i < imgs.size(); i++) {
      if (out[i].status == ZW_OK && out[i].data) res[i].assign(out[i].data, out[i].data + out[i].len);
      else if (!first_err) first_err = out[i].status;
      zw_free(out[i].data);
    }
    raise_for(first_err);
    return res;
  }
 private:
  zw_multi* h_;
};

// WebPEncoder::new(&mut Vec<u8>) / set_params / encode: appends the .webp bytes to `writer`.
class WebPEncoder {
 public:
  explicit WebPEncoder(std::vector<uint8_t>& writer, Context& ctx) : w_(writer), ctx_(ctx) {}
  void set_params(const EncoderParams& p) { params_ = p; }
  void encode(const uint8_t* data, size_t len, uint32_t width, uint32_t height, ColorType color) {
    if (width > 65535 || height > 65535) throw InvalidDimensions();
    auto out = ctx_.encode_batch({Context::ImageRef{data, len, width, height, color}}, params_);
    w_.insert(w_.end(), out[0].begin(), out[0].end());
  }
 private:
  std::vector<uint8_t>& w_;
  Context& ctx_;
  EncoderParams params_;
};

}  // namespace zenwebp_b200
#endif
