"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, fails loudly without a GPU (no CPU fallback), API mirror semantics, sharding."""
import ctypes as C
import os
import re

import pytest

import oracle_lib as O

ROOT = O.ROOT


def _lib():
    from image_webp_b200 import _lib
    return _lib


def test_library_exports_every_header_symbol():
    L = _lib().load()
    hdr = open(os.path.join(ROOT, "include", "zenwebp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(zw_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 12
    assert sorted(names) == sorted(_lib().EXPORTS)
    for n in names:
        assert hasattr(L, n), "library does not export %s" % n
    assert b"sm_100a" in L.zw_version()


def test_strerror_and_max_output():
    L = _lib().load()
    assert L.zw_strerror(0) == b"ok"
    assert b"dimension" in L.zw_strerror(1)
    assert b"buffer" in L.zw_strerror(2)
    assert L.zw_max_output_size(768, 512) > 768 * 512


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import image_webp_b200 as Z
    with pytest.raises(Z.DeviceError) as ei:
        Z.Context(0)
    assert "no CPU fallback" in str(ei.value)
    out = bytearray()
    enc = Z.WebPEncoder(out)
    enc.set_params(Z.EncoderParams.lossy(75))
    with pytest.raises(Z.DeviceError):
        enc.encode(bytes(16 * 16 * 3), 16, 16, Z.ColorType.Rgb8)
    assert out == bytearray()


def test_product_never_touches_the_oracle():
    # the product path must not import / link / include anything under oracle/ or tests/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "image_webp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt, f
                assert "zw_oracle" not in txt and "oracle_lib" not in txt and "zwo_" not in txt, f
    for f in ("include/zenwebp_b200.h",):
        assert "zwo_" not in open(os.path.join(ROOT, f)).read()


def test_encoder_params_mirror_reference_defaults():
    import image_webp_b200 as Z
    p = Z.EncoderParams()                 # api.rs:434-443
    assert (p.use_predictor_transform, p.use_lossy, p.lossy_quality, p.method) == (True, False, 95, 4)
    q = Z.EncoderParams.lossy(75)         # api.rs:452-458
    assert q.use_lossy and q.lossy_quality == 75 and q.method == 4
    assert Z.EncoderParams.lossy(10).with_method(6).method == 6
    assert Z.ColorType.Rgb8.bytes_per_pixel() == 3 and Z.ColorType.Rgba8.has_alpha() and not Z.ColorType.L8.has_alpha()


def test_builder_api_mirrors_reference():
    import image_webp_b200 as Z
    c = Z.EncoderConfig.new()             # api.rs:501-513: lossy, quality 75, method 4
    assert (c.get_quality(), c.get_preset(), c.is_lossless(), c.get_method()) == (75.0, Z.Preset.Default, False, 4)
    assert Z.EncoderConfig.new_lossless().is_lossless()
    assert c.quality(120).get_quality() == 100.0 and c.quality(-3).get_quality() == 0.0   # clamp, api.rs:548
    assert c.method(9).get_method() == 6                                                   # min(6), api.rs:569
    assert Z.EncoderConfig.with_preset(Z.Preset.Photo, 85).get_preset() == Z.Preset.Photo
    for q, want in ((75.4, 75), (75.5, 76), (0.0, 0), (0.4, 0), (0.5, 1), (100.0, 100)):   # fast_math::roundf, fast_math.rs:129-137
        p = Z.EncoderConfig().quality(q).to_params()
        assert p.lossy_quality == want and p.use_lossy and p.use_predictor_transform
    assert not Z.EncoderConfig().lossless(True).to_params().use_lossy
    e = Z.Encoder.new_rgb(b"", 4, 4).quality(50).preset(Z.Preset.Text).near_lossless(200).alpha_quality(7).exact(True).target_size(9).sharp_yuv(True)
    assert e._config._near_lossless == 100 and e._config._alpha_quality == 7 and e._config._target_size == 9
    with pytest.raises(Z.InvalidBufferSize):   # validate_buffer_size runs before any device work (api.rs:871-876)
        e.encode()


def test_shard_ranges_cover_and_are_contiguous():
    from image_webp_b200 import shard
    for n in (0, 1, 7, 8, 9, 1024, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            rs = [shard.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            for a, b in zip(rs, rs[1:]):
                assert a[1] == b[0]
            assert max(e - b for b, e in rs) == (n + w - 1) // w
            assert sum(shard.shard_sizes(n, w)) == n
    assert shard.gather_in_order([[1, 2], [], [3]]) == [1, 2, 3]
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)
