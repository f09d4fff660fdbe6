// zenwebp_b200.hpp -- C++ host-side mirror of the reference's encoder API (lossy VP8, lossless VP8L, every
// container), header-only on top of the C ABI (zenwebp_b200.h).  Mirrors imazen/image-webp `zenwebp` 0.2.0:
//   ColorType      src/encoder/api.rs:83-92      EncodingError  src/encoder/api.rs:35-48
//   EncoderParams  src/encoder/api.rs:419-459    WebPEncoder    src/encoder/api.rs:1244-1398
//   EncoderConfig  src/encoder/api.rs:488-672    Encoder        src/encoder/api.rs:703-914
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
  struct Metadata { std::vector<uint8_t> icc_profile, exif, xmp; };
  // Batch entry: n images -> n .webp files exactly as WebPEncoder::encode writes them (lossy or lossless frame, simple
  // or extended container; `meta`: empty or one entry per image).  Per-image errors throw for the first failing image.
  std::vector<std::vector<uint8_t>> encode_batch(const std::vector<ImageRef>& imgs, const EncoderParams& p, zw_timing* t = nullptr,
                                                 const std::vector<Metadata>& meta = {}) {
    if (!meta.empty() && meta.size() != imgs.size()) throw std::logic_error("one metadata entry per image");
    std::vector<zw_image> in(imgs.size());
    std::vector<zw_output> out(imgs.size());
    std::vector<zw_metadata> md(meta.size());
    for (size_t i = 0; i < imgs.size(); i++) {
      in[i] = zw_image{imgs[i].data, imgs[i].len, imgs[i].width, imgs[i].height, (uint32_t)imgs[i].color, 0};
      out[i] = zw_output{nullptr, 0, 0, 0, 0};
    }
    for (size_t i = 0; i < meta.size(); i++)
      md[i] = zw_metadata{meta[i].icc_profile.data(), meta[i].icc_profile.size(), meta[i].exif.data(), meta[i].exif.size(),
                          meta[i].xmp.data(), meta[i].xmp.size()};
    const zw_params zp{p.use_predictor_transform ? 1 : 0, p.use_lossy ? 1 : 0, p.lossy_quality, p.method};
    const int rc = zw_encode_batch(h_, in.data(), in.size(), &zp, md.empty() ? nullptr : md.data(), out.data(), t);
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

// WebPEncoder::new(&mut Vec<u8>) / set_params / set_*_metadata / encode: appends the .webp bytes to `writer`.
class WebPEncoder {
 public:
  explicit WebPEncoder(std::vector<uint8_t>& writer, Context& ctx) : w_(writer), ctx_(ctx) {}
  void set_params(const EncoderParams& p) { params_ = p; }
  void set_icc_profile(std::vector<uint8_t> v) { meta_.icc_profile = std::move(v); }
  void set_exif_metadata(std::vector<uint8_t> v) { meta_.exif = std::move(v); }
  void set_xmp_metadata(std::vector<uint8_t> v) { meta_.xmp = std::move(v); }
  void encode(const uint8_t* data, size_t len, uint32_t width, uint32_t height, ColorType color) {
    const bool any = !meta_.icc_profile.empty() || !meta_.exif.empty() || !meta_.xmp.empty();
    auto out = ctx_.encode_batch({Context::ImageRef{data, len, width, height, color}}, params_, nullptr,
                                 any ? std::vector<Context::Metadata>{meta_} : std::vector<Context::Metadata>{});
    w_.insert(w_.end(), out[0].begin(), out[0].end());
  }
 private:
  std::vector<uint8_t>& w_;
  Context& ctx_;
  EncoderParams params_;
  Context::Metadata meta_;
};

enum class Preset { Default, Picture, Photo, Drawing, Icon, Text };  // api.rs:54-76

// EncoderConfig (api.rs:488-672): reusable builder.  Default: lossy, quality 75, method 4.  The knobs the reference
// stores without reading (preset, near_lossless, alpha_quality, exact, target_size, sharp_yuv) are stored here too.
class EncoderConfig {
 public:
  static EncoderConfig new_lossless() { EncoderConfig c; c.lossless_ = true; return c; }
  static EncoderConfig with_preset(Preset p, float q) { EncoderConfig c; c.preset_ = p; c.quality_ = q; return c; }
  EncoderConfig& quality(float q) { quality_ = q < 0.f ? 0.f : (q > 100.f ? 100.f : q); return *this; }
  EncoderConfig& preset(Preset p) { preset_ = p; return *this; }
  EncoderConfig& lossless(bool v) { lossless_ = v; return *this; }
  EncoderConfig& method(uint8_t m) { method_ = m < 6 ? m : 6; return *this; }
  EncoderConfig& near_lossless(uint8_t v) { near_lossless_ = v < 100 ? v : 100; return *this; }
  EncoderConfig& alpha_quality(uint8_t v) { alpha_quality_ = v < 100 ? v : 100; return *this; }
  EncoderConfig& exact(bool v) { exact_ = v; return *this; }
  EncoderConfig& target_size(uint32_t v) { target_size_ = v; return *this; }
  EncoderConfig& sharp_yuv(bool v) { sharp_yuv_ = v; return *this; }
  float get_quality() const { return quality_; }
  Preset get_preset() const { return preset_; }
  bool is_lossless() const { return lossless_; }
  uint8_t get_method() const { return method_; }
  EncoderParams to_params() const {  // api.rs:633-640; fast_math::roundf = (x + 0.5) as i32
    EncoderParams p;
    p.use_predictor_transform = true; p.use_lossy = !lossless_; p.method = method_;
    const int q = (int)(quality_ + 0.5f);
    p.lossy_quality = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
    return p;
  }
  std::vector<uint8_t> encode(Context& ctx, const uint8_t* data, size_t len, uint32_t w, uint32_t h, ColorType color,
                              const Context::Metadata* meta = nullptr) const {
    if ((uint64_t)len < (uint64_t)w * h * ((uint32_t)color + 1)) throw InvalidBufferSize();  // validate_buffer_size, api.rs:917
    std::vector<uint8_t> out;
    WebPEncoder enc(out, ctx);
    enc.set_params(to_params());
    if (meta) { enc.set_icc_profile(meta->icc_profile); enc.set_exif_metadata(meta->exif); enc.set_xmp_metadata(meta->xmp); }
    enc.encode(data, len, w, h, color);
    return out;
  }
  std::vector<uint8_t> encode_rgba(Context& ctx, const uint8_t* d, size_t n, uint32_t w, uint32_t h) const { return encode(ctx, d, n, w, h, ColorType::Rgba8); }
  std::vector<uint8_t> encode_rgb(Context& ctx, const uint8_t* d, size_t n, uint32_t w, uint32_t h) const { return encode(ctx, d, n, w, h, ColorType::Rgb8); }
 private:
  float quality_ = 75.f;
  Preset preset_ = Preset::Default;
  bool lossless_ = false, exact_ = false, sharp_yuv_ = false;
  uint8_t method_ = 4, near_lossless_ = 100, alpha_quality_ = 100;
  uint32_t target_size_ = 0;
};

// Encoder (api.rs:703-914): Encoder::new_rgba(data, w, h).quality(85).encode(ctx).
class Encoder {
 public:
  static Encoder new_rgba(const uint8_t* d, size_t n, uint32_t w, uint32_t h) { return Encoder(d, n, w, h, ColorType::Rgba8); }
  static Encoder new_rgb(const uint8_t* d, size_t n, uint32_t w, uint32_t h) { return Encoder(d, n, w, h, ColorType::Rgb8); }
  static Encoder new_l8(const uint8_t* d, size_t n, uint32_t w, uint32_t h) { return Encoder(d, n, w, h, ColorType::L8); }
  static Encoder new_la8(const uint8_t* d, size_t n, uint32_t w, uint32_t h) { return Encoder(d, n, w, h, ColorType::La8); }
  Encoder& quality(float q) { cfg_.quality(q); return *this; }
  Encoder& preset(Preset p) { cfg_.preset(p); return *this; }
  Encoder& lossless(bool v) { cfg_.lossless(v); return *this; }
  Encoder& method(uint8_t m) { cfg_.method(m); return *this; }
  Encoder& config(const EncoderConfig& c) { cfg_ = c; return *this; }
  Encoder& icc_profile(std::vector<uint8_t> v) { meta_.icc_profile = std::move(v); return *this; }
  Encoder& exif_metadata(std::vector<uint8_t> v) { meta_.exif = std::move(v); return *this; }
  Encoder& xmp_metadata(std::vector<uint8_t> v) { meta_.xmp = std::move(v); return *this; }
  std::vector<uint8_t> encode(Context& ctx) const { return cfg_.encode(ctx, d_, n_, w_, h_, color_, &meta_); }
  void encode_into(Context& ctx, std::vector<uint8_t>& out) const { auto e = encode(ctx); out.insert(out.end(), e.begin(), e.end()); }
 private:
  Encoder(const uint8_t* d, size_t n, uint32_t w, uint32_t h, ColorType c) : d_(d), n_(n), w_(w), h_(h), color_(c) {}
  const uint8_t* d_; size_t n_; uint32_t w_, h_; ColorType color_;
  EncoderConfig cfg_;
  Context::Metadata meta_;
};

}  // namespace zenwebp_b200
#endif
