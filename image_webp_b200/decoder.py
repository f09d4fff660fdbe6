"""Host-side mirror of the reference's lossy DECODE path on top of the C ABI (zw_decode_batch / zw_verify): the
on-device VP8 key-frame decoder used as the batch verifier.

Mirrors of imazen/image-webp `zenwebp` 0.2.0 (names, argument meaning, error behaviour):
  UpsamplingMethod    src/decoder/api.rs:268       (Bilinear = default, Simple)
  DecodingError       src/decoder/api.rs:7         (the variants this path can raise)
  WebPDecoder         src/decoder/api.rs:283-905   (new / dimensions / has_alpha / is_lossy / is_animated /
                                                    output_buffer_size / set_lossy_upsampling / read_image)
  decode_rgb          src/decoder/api.rs:973
plus the batch entries `decode_batch` / `verify_batch`.  Only lossy still images (a 'VP8 ' chunk) are this path:
lossless (VP8L) and animated files raise UnsupportedFeature; the alpha plane of VP8X files is not decoded (colour
only).  No CPU fallback: everything after the container scan runs in CUDA kernels."""
import ctypes as C
import enum

import numpy as np

from . import _lib
from .encoder import ColorType, DeviceError, default_context


class UpsamplingMethod(enum.Enum):
    Bilinear = 1
    Simple = 0


class DecodingError(Exception):
    KINDS = {1: "BitStreamError", 2: "UnsupportedFeature(Non-keyframe frames)", 3: "Vp8MagicInvalid", 4: "ColorSpaceInvalid",
             5: "BitStreamError (truncated)", 6: "UnsupportedFeature(no lossy 'VP8 ' chunk)", 7: "InconsistentImageSizes"}

    def __init__(self, code):
        super().__init__(self.KINDS.get(code, "decoder error %d" % code))
        self.code = code


def _info_dict(i):
    return i.as_dict()


def decode_batch(files, upsampling=UpsamplingMethod.Bilinear, sources=None, color=ColorType.Rgb8, want_rgb=True, ctx=None,
                 raise_errors=True):
    """Decode a list of .webp files / bare VP8 frames (bytes) on the GPU.  Returns (list of uint8 arrays [h, w, 3] or
    None, list of info dicts, (parse_ms, reconstruct_ms, filter_ms, colour_ms)).  With `sources` (arrays [h, w, c] matching `color`) every image is
    also scored against its source on the device: info['sse_rgb'], info['psnr_rgb']."""
    ctx = ctx or default_context()
    n = len(files)
    blobs = (_lib.ZwBlob * n)()
    keep = []
    for i, f in enumerate(files):
        b = np.frombuffer(bytes(f), dtype=np.uint8)
        keep.append(b)
        blobs[i] = _lib.ZwBlob(b.ctypes.data if b.size else None, b.size)
    outs = (_lib.ZwOutput * n)() if want_rgb else None
    src_arr = None
    if sources is not None:
        src_arr, keep2 = ctx._as_images(sources, color)
        keep.append(keep2)
    infos = (_lib.ZwDecodeInfo * n)()
    ms = (C.c_float * 4)()
    rc = ctx.lib.zw_decode_batch(ctx.h, blobs, n, upsampling.value, outs, src_arr, infos, ms)
    res = []
    if want_rgb:
        for i in range(n):
            o = outs[i]
            if o.status == 0 and infos[i].status == 0:
                res.append(np.frombuffer(C.string_at(o.data, o.len), np.uint8).reshape(infos[i].height, infos[i].width, 3))
            else:
                res.append(None)
            if o.data:
                ctx.lib.zw_free(o.data)
    else:
        res = [None] * n
    if rc != 0:
        raise DeviceError(rc, ctx.lib.zw_strerror(rc).decode())
    info = [_info_dict(infos[i]) for i in range(n)]
    if raise_errors:
        for d in info:
            if d["status"] != 0:
                raise DecodingError(d["status"])
    return res, info, tuple(ms)


def verify_pending(pending, upsampling=UpsamplingMethod.Bilinear):
    """Decode the files of a submitted batch (encoder.PendingBatch, before .result() releases it) where they lie in
    device memory and score them against the batch's source pixels, which are still resident there too (zw_verify).
    Returns (list of info dicts, (parse_ms, reconstruct_ms, filter_ms, colour_ms))."""
    ctx = pending._ctx
    n = pending._n
    infos = (_lib.ZwDecodeInfo * n)()
    ms = (C.c_float * 4)()
    rc = ctx.lib.zw_verify(ctx.h, pending._ticket, upsampling.value, infos, ms)
    if rc != 0:
        raise DeviceError(rc, ctx.lib.zw_strerror(rc).decode())
    return [_info_dict(infos[i]) for i in range(n)], tuple(ms)


def _scan(data):
    """(kind, width, height, has_alpha, animated) from the container: kind 'VP8 ' / 'VP8L' / None."""
    d = bytes(data)
    if d[:4] != b"RIFF" or d[8:12] != b"WEBP":
        if len(d) >= 10 and d[3:6] == b"\x9d\x01\x2a":
            return "VP8 ", (d[6] | d[7] << 8) & 0x3FFF, (d[8] | d[9] << 8) & 0x3FFF, False, False
        return None, 0, 0, False, False
    pos, kind, w, h, alpha, anim = 12, None, 0, 0, False, False
    while pos + 8 <= len(d):
        cc, sz = d[pos:pos + 4], int.from_bytes(d[pos + 4:pos + 8], "little")
        body = d[pos + 8:pos + 8 + sz]
        if cc == b"VP8X" and sz >= 10:
            alpha, anim = bool(body[0] & 0x10), bool(body[0] & 0x02)
        elif cc == b"ALPH":
            alpha = True
        elif cc == b"VP8 " and kind is None and sz >= 10:
            kind, w, h = "VP8 ", (body[6] | body[7] << 8) & 0x3FFF, (body[8] | body[9] << 8) & 0x3FFF
        elif cc == b"VP8L" and kind is None:
            kind = "VP8L"
        pos += 8 + sz + (sz & 1)
    return kind, w, h, alpha, anim


class WebPDecoder:
    """WebPDecoder (src/decoder/api.rs:283) for lossy still images, decoding on the GPU."""

    def __init__(self, data, ctx=None):
        self._data = bytes(data)
        self._ctx = ctx
        self._kind, self._w, self._h, self._alpha, self._anim = _scan(self._data)
        if self._kind is None:
            raise DecodingError(6)
        self._upsampling = UpsamplingMethod.Bilinear

    def dimensions(self):
        return (self._w, self._h)

    def has_alpha(self):
        return self._alpha

    def is_animated(self):
        return self._anim

    def is_lossy(self):
        return self._kind == "VP8 "

    def output_buffer_size(self):
        return self._w * self._h * (4 if self._alpha else 3)

    def set_lossy_upsampling(self, upsampling_method):
        self._upsampling = upsampling_method

    def read_image(self):
        """RGB pixels [h, w, 3] (the reference fills a caller buffer; alpha of VP8X files is not decoded here)."""
        if self._kind != "VP8 " or self._anim:
            raise DecodingError(6)
        out, _, _ = decode_batch([self._data], self._upsampling, ctx=self._ctx)
        return out[0]


def decode_rgb(data):
    """decode_rgb (src/decoder/api.rs:973): (pixels [h, w, 3], width, height)."""
    d = WebPDecoder(data)
    px = d.read_image()
    return px, px.shape[1], px.shape[0]
