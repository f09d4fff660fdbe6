// hpp_driver.cpp -- TEST INFRASTRUCTURE: compiles include/zenwebp_b200.hpp (the C++ mirror of the reference's encoder API)
// against the C ABI and drives one encode through every class of it.  `hpp_driver selftest` needs no GPU (it checks the
// error path of a context that cannot be created); `hpp_driver encode <w> <h> <rgb-file> <out-prefix>` encodes the same
// image through Context::encode_batch, WebPEncoder, BatchPipeline and MultiContext and writes the four files.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "../../include/zenwebp_b200.hpp"
namespace Z = zenwebp_b200;

static void write_file(const std::string& path, const std::vector<uint8_t>& d) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(d.data()), (std::streamsize)d.size());
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "selftest")) {
    Z::EncoderParams p = Z::EncoderParams::lossy(75).with_method(4);
    if (!(p.use_lossy && p.lossy_quality == 75 && p.method == 4 && !Z::EncoderParams().use_lossy && Z::EncoderParams().lossy_quality == 95)) return 2;
    try {
      Z::Context c(1 << 20);  // no such device (or no driver at all): must throw, never fall back
      return 3;
    } catch (const Z::EncodingError& e) {
      printf("selftest ok: %s\n", e.what());
      return 0;
    }
  }
  if (argc == 6 && !strcmp(argv[1], "encode")) {
    const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]);
    std::ifstream f(argv[4], std::ios::binary);
    std::vector<uint8_t> rgb((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const std::string prefix = argv[5];
    const Z::EncoderParams p = Z::EncoderParams::lossy(75).with_method(4);
    const Z::Context::ImageRef ref{rgb.data(), rgb.size(), w, h, Z::ColorType::Rgb8};
    Z::Context ctx(0);
    write_file(prefix + ".batch.webp", ctx.encode_batch({ref, ref}, p)[1]);
    std::vector<uint8_t> out = {'x'};
    Z::WebPEncoder enc(out, ctx);
    enc.set_params(p);
    enc.encode(rgb.data(), rgb.size(), w, h, Z::ColorType::Rgb8);
    write_file(prefix + ".encoder.webp", std::vector<uint8_t>(out.begin() + 1, out.end()));
    {
      Z::BatchPipeline pipe(0, 2);
      auto a = pipe.submit({ref}, p);
      auto b = pipe.submit({ref, ref}, p);
      bool busy = false;
      try { pipe.submit({ref}, p); } catch (const Z::EncodingError& e) { busy = e.code == ZW_ERR_BUSY; }
      if (!busy) return 4;
      write_file(prefix + ".pipe.webp", a.get()[0]);
      if (b.get().size() != 2) return 5;
    }
    Z::MultiContext mc({0, 0});
    write_file(prefix + ".multi.webp", mc.encode_batch({ref, ref, ref}, p)[2]);
    try { ctx.encode_batch({Z::Context::ImageRef{rgb.data(), rgb.size() - 1, w, h, Z::ColorType::Rgb8}}, p); return 6; }
    catch (const Z::InvalidBufferSize&) {}
    {  // decoder mirror: decode what was just encoded, scored against the source; the pixels go to <prefix>.decoded.rgb
      const std::vector<uint8_t> file = ctx.encode_batch({ref}, p)[0];
      const std::vector<Z::Context::ImageRef> srcs = {ref};
      auto dec = ctx.decode_batch({{file.data(), file.size()}}, false, &srcs);
      if (dec.size() != 1 || dec[0].width != w || dec[0].height != h || dec[0].rgb.size() != (size_t)w * h * 3 || dec[0].psnr_rgb < 20.0) return 7;
      write_file(prefix + ".decoded.rgb", dec[0].rgb);
      try { ctx.decode_batch({{file.data(), 5}}); return 8; } catch (const std::runtime_error&) {}
    }
    {  // lossless + builder + metadata: <prefix>.lossless.webp (Encoder, default metadata-free), <prefix>.meta.webp (lossy + EXIF)
      write_file(prefix + ".lossless.webp", Z::Encoder::new_rgb(rgb.data(), rgb.size(), w, h).lossless(true).encode(ctx));
      write_file(prefix + ".meta.webp", Z::Encoder::new_rgb(rgb.data(), rgb.size(), w, h).quality(74.6f).method(9).exif_metadata({'E', 'X', 'I'}).encode(ctx));
      Z::EncoderConfig cfg = Z::EncoderConfig::new_lossless();
      if (!cfg.is_lossless() || cfg.to_params().use_lossy || Z::EncoderConfig().quality(75.5f).to_params().lossy_quality != 76) return 9;
      try { Z::Encoder::new_rgb(rgb.data(), 5, w, h).encode(ctx); return 10; } catch (const Z::InvalidBufferSize&) {}
    }
    printf("encode ok\n");
    return 0;
  }
  fprintf(stderr, "usage: hpp_driver selftest | encode <w> <h> <rgb-file> <out-prefix>\n");
  return 1;
}
