"""Host-side mirror of the reference's encoder API, on top of the C ABI.

Mirrors (names, argument meaning, error behaviour) of imazen/image-webp `zenwebp` 0.2.0:
  ColorType        src/encoder/api.rs:83-92
  EncodingError    src/encoder/api.rs:35-48
  Preset           src/encoder/api.rs:54-76
  EncoderParams    src/encoder/api.rs:419-459   (+ additive .method(m) builder)
  EncoderConfig    src/encoder/api.rs:488-672   (builder; the knobs the reference stores but never reads are stored too)
  Encoder          src/encoder/api.rs:703-914   (new_rgba / new_rgb / new_l8 / new_la8 ... encode / encode_into / encode_to_writer)
  WebPEncoder      src/encoder/api.rs:1244-1398 (new / set_params / set_icc_profile / set_exif_metadata / set_xmp_metadata / encode)
plus the batch entry point north_star asks for (`encode_batch`).  Lossy (VP8) and lossless (VP8L) frames, the simple and
the extended (VP8X + ICCP + ALPH + EXIF + XMP) container: everything WebPEncoder::encode writes.

The reference is Rust; no Rust toolchain exists in the build image, so this module (and the C++
header include/zenwebp_b200.hpp) stand where the `zenwebp-b200` wrapper crate would: see
INTEGRATION.md for the -sys crate a maintainer would add.  No CPU fallback.
"""
import ctypes as C
import enum
import threading

import numpy as np

from . import _lib


class ColorType(enum.Enum):
    L8 = 0
    La8 = 1
    Rgb8 = 2
    Rgba8 = 3

    def bytes_per_pixel(self):
        return {ColorType.L8: 1, ColorType.La8: 2, ColorType.Rgb8: 3, ColorType.Rgba8: 4}[self]

    def has_alpha(self):
        return self in (ColorType.La8, ColorType.Rgba8)


class Preset(enum.Enum):
    """api.rs:54-76.  Stored by the builder; like the reference, the encoder does not read it."""
    Default = 0
    Picture = 1
    Photo = 2
    Drawing = 3
    Icon = 4
    Text = 5


class EncodingError(Exception):
    """EncodingError::{InvalidDimensions, InvalidBufferSize(String)} (+ device errors)."""

    def __init__(self, kind, message=""):
        super().__init__(("%s: %s" % (kind, message)) if message else kind)
        self.kind = kind


class InvalidDimensions(EncodingError):
    def __init__(self):
        super().__init__("Invalid dimensions")


class InvalidBufferSize(EncodingError):
    def __init__(self, msg=""):
        super().__init__("Invalid buffer size", msg)


class DeviceError(EncodingError):
    def __init__(self, code, msg):
        super().__init__("CUDA/encoder error %d" % code, msg)
        self.code = code


class EncoderParams:
    """api.rs:419-459.  Public fields like the reference; Default = lossless, q95, method 4."""

    def __init__(self, use_predictor_transform=True, use_lossy=False, lossy_quality=95, method=4):
        self.use_predictor_transform = use_predictor_transform
        self.use_lossy = use_lossy
        self.lossy_quality = lossy_quality
        self.method = method

    @classmethod
    def lossless(cls):
        return cls()

    @classmethod
    def lossy(cls, quality):
        return cls(use_lossy=True, lossy_quality=quality)

    def with_method(self, m):  # the additive `EncoderParams::method(self, m) -> Self`
        self.method = m
        return self


def _raise_for(status, lib):
    if status == 0:
        return
    if status == 1:
        raise InvalidDimensions()
    if status == 2:
        raise InvalidBufferSize("width/height doesn't match data length")
    if status == 3:
        raise ValueError("lossy quality must be between 0 and 100 / unsupported parameter")
    raise DeviceError(status, lib.zw_strerror(status).decode())


class PendingBatch:
    """A batch in flight (zw_submit ticket).  `result()` blocks until it is finished and returns
    (list of .webp / VP8 bytes, timing dict); the pipeline slot is released then."""

    def __init__(self, ctx, ticket, keep, container, raise_errors, views=False):
        self._views = views
        self._ctx, self._ticket, self._keep = ctx, ticket, keep
        self._n = len(keep[0])
        self._container, self._raise = container, raise_errors
        self._res = None

    def result(self, views=None):
        """views=True (or a batch submitted with views=True): the files come back as memoryviews into ONE copy of the slot's pinned arena (a single memcpy per
        batch instead of one bytes object per image); they compare equal to bytes and go wherever a buffer is accepted."""
        if self._res is None:
            ctx = self._ctx
            view = _lib.ZwBatchView()
            t = _lib.ZwTiming()
            rc = ctx.lib.zw_wait(ctx.h, self._ticket, 1 if self._container else 0, C.byref(view), C.byref(t))
            if rc != 0:
                ctx.lib.zw_release(ctx.h, self._ticket)
                _raise_for(rc, ctx.lib)
            outs, first_bad = [], 0
            n = view.n
            offs = np.ctypeslib.as_array(view.offsets, (n,)) if n else np.zeros(0, np.uint64)
            lens = np.ctypeslib.as_array(view.lens, (n,)) if n else np.zeros(0, np.uint32)
            stat = np.ctypeslib.as_array(view.status, (n,)) if n else np.zeros(0, np.int32)
            blob, lo = None, 0
            views = self._views if views is None else views
            if views and n and (stat == 0).any():
                ok = stat == 0
                lo = int(offs[ok].min())
                hi = int((offs[ok] + lens[ok]).max())
                blob = memoryview(C.string_at(view.arena + lo, hi - lo))
            for i in range(n):
                st = int(stat[i])
                if st != 0:
                    first_bad = first_bad or st
                    outs.append(None)
                elif blob is not None:
                    o = int(offs[i]) - lo
                    outs.append(blob[o:o + int(lens[i])])
                else:
                    outs.append(C.string_at(view.arena + int(offs[i]), int(lens[i])))
            ctx.lib.zw_release(ctx.h, self._ticket)
            self._keep = None
            self._res = (outs, t.as_dict())
            if first_bad and self._raise:
                _raise_for(first_bad, ctx.lib)
        return self._res


class PreparedBatch:
    """zw_image descriptors of a batch (Context.prepare)."""

    def __init__(self, arr, keep, color):
        self.arr, self.keep, self.color = arr, keep, color

    def __len__(self):
        return len(self.keep)


class Context:
    """One encoder context per (host thread, GPU) -- wraps zw_create / zw_destroy."""

    def __init__(self, device=0, max_device_bytes=0, persistent_warps_per_sm=0, depth=0, lanes=0):
        self.lib = _lib.load()
        res = (C.c_int * 5)()
        res[0] = depth or lanes  # batches the context keeps in flight (0 = library default)
        lim = _lib.ZwLimits(max_device_bytes, persistent_warps_per_sm, res)
        self.h = self.lib.zw_create(device, C.byref(lim))
        if not self.h:
            code = self.lib.zw_last_error()
            raise DeviceError(code, self.lib.zw_strerror(code).decode() + " (no CPU fallback exists)")
        self.device = device
        self._keep = None

    def close(self):
        if self.h:
            self.lib.zw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------
    @staticmethod
    def prepare(images, color=ColorType.Rgb8):
        """Build the zw_image descriptors of a batch once; the result can be handed to encode_batch / submit any number
        of times (the pixel buffers are referenced, not copied: keep them unchanged while a batch is in flight)."""
        return PreparedBatch(*Context._as_images(images, color), color)

    @staticmethod
    def _as_images(images, color):
        if isinstance(images, PreparedBatch):
            if images.color != color:
                raise ValueError("batch was prepared for %s" % images.color)
            return images.arr, images.keep
        arr = (_lib.ZwImage * len(images))()
        keep = []
        for i, im in enumerate(images):
            if isinstance(im, tuple):  # (bytes-like, width, height)
                data, w, h = im
                buf = np.frombuffer(data, dtype=np.uint8)
            else:
                buf = np.ascontiguousarray(im, dtype=np.uint8)
                h, w = buf.shape[0], buf.shape[1]
            keep.append(buf)
            arr[i] = _lib.ZwImage(buf.ctypes.data if buf.size else None, buf.size, w, h, color.value, 0)
        return arr, keep

    def _collect(self, outs, raise_errors):
        """Copy every output out of its malloc'ed buffer and free it; raise (if asked) only afterwards."""
        res, first_bad = [], 0
        for o in outs:
            if o.status != 0:
                first_bad = first_bad or o.status
                res.append(None)
            else:
                res.append(C.string_at(o.data, o.len))
            if o.data:
                self.lib.zw_free(o.data)
        if first_bad and raise_errors:
            _raise_for(first_bad, self.lib)
        return res

    @staticmethod
    def _check(params, color, container):
        """The streaming / multi-GPU entries carry the lossy path with the simple container only."""
        if not params.use_lossy:
            raise NotImplementedError("streaming entry: lossy only (Context.encode_batch handles lossless)")
        if container and color.has_alpha():
            raise NotImplementedError("streaming entry: lossy+alpha needs VP8X + ALPH (Context.encode_batch builds it)")

    @staticmethod
    def _as_metadata(metadata, n):
        """metadata: None, one dict(icc=, exif=, xmp=) for every image, or a list of n dicts / None."""
        if metadata is None:
            return None, None
        items = [metadata] * n if isinstance(metadata, dict) else list(metadata)
        if len(items) != n:
            raise ValueError("metadata: one entry per image")
        arr = (_lib.ZwMetadata * n)()
        keep = []
        for i, m in enumerate(items):
            m = m or {}
            icc, exif, xmp = bytes(m.get("icc", b"")), bytes(m.get("exif", b"")), bytes(m.get("xmp", b""))
            keep.append((icc, exif, xmp))
            arr[i] = _lib.ZwMetadata(icc or None, len(icc), exif or None, len(exif), xmp or None, len(xmp))
        return arr, keep

    # -- public ----------------------------------------------------------------------------
    def encode_batch(self, images, params, color=ColorType.Rgb8, container=True, raise_errors=True, metadata=None):
        """Batch entry point: list of uint8 arrays [h,w,c] (or (bytes,w,h) tuples) -> list of bytes.
        Returns (outputs, timing dict).  container=True: complete .webp files exactly as WebPEncoder::encode writes them
        (zw_encode_batch: lossy or lossless, simple or extended container, optional metadata per image);
        container=False: the raw frame (VP8 payload, zw_encode_vp8_batch, or VP8L stream, zw_encode_lossless_batch)."""
        arr, keep = self._as_images(images, color)
        n = len(images)
        outs = (_lib.ZwOutput * n)()
        t = _lib.ZwTiming()
        if container:
            zp = _lib.ZwParams(1 if params.use_predictor_transform else 0, 1 if params.use_lossy else 0, int(params.lossy_quality), int(params.method))
            marr, mkeep = self._as_metadata(metadata, n)
            rc = self.lib.zw_encode_batch(self.h, arr, n, C.byref(zp), marr, outs, C.byref(t))
        elif metadata is not None:
            raise ValueError("metadata needs a container")
        elif params.use_lossy:
            rc = self.lib.zw_encode_vp8_batch(self.h, arr, n, int(params.lossy_quality), int(params.method), outs, C.byref(t))
        else:
            rc = self.lib.zw_encode_lossless_batch(self.h, arr, n, 1 if params.use_predictor_transform else 0, 0, outs, C.byref(t))
        if rc != 0:
            self._collect(outs, False)
            _raise_for(rc, self.lib)
        return self._collect(outs, raise_errors), t.as_dict()

    def encode_batch_raw(self, images, params, color=ColorType.Rgb8):
        """zw_encode_webp_batch with library-allocated outputs left in place: returns (zw_output array, return code).  The
        caller hands the array to `_collect` (copies the files out and frees them).  For timing the C call alone."""
        self._check(params, color, True)
        arr, keep = self._as_images(images, color)
        outs = (_lib.ZwOutput * len(keep))()
        rc = self.lib.zw_encode_webp_batch(self.h, arr, len(keep), int(params.lossy_quality), int(params.method), outs, None)
        return outs, rc

    def encode_alpha_batch(self, images, color=ColorType.Rgba8, raise_errors=True):
        """ALPH chunk payloads (encode_alpha_lossless, api.rs:1175) of La8 / Rgba8 images."""
        arr, keep = self._as_images(images, color)
        outs = (_lib.ZwOutput * len(images))()
        t = _lib.ZwTiming()
        rc = self.lib.zw_encode_alpha_batch(self.h, arr, len(images), outs, C.byref(t))
        if rc != 0:
            self._collect(outs, False)
            _raise_for(rc, self.lib)
        return self._collect(outs, raise_errors), t.as_dict()

    def lossless_dump_stage(self, index, name, dtype=np.uint8):
        n = C.c_size_t(0)
        rc = self.lib.zw_lossless_dump_stage(self.h, index, name.encode(), None, 0, C.byref(n))
        if rc not in (0, 4):
            _raise_for(rc, self.lib)
        buf = np.zeros(max(1, n.value), np.uint8)
        rc = self.lib.zw_lossless_dump_stage(self.h, index, name.encode(), buf.ctypes.data, buf.size, C.byref(n))
        if rc != 0:
            _raise_for(rc, self.lib)
        return buf[:n.value].view(dtype)

    def submit(self, images, params, color=ColorType.Rgb8, container=True, raise_errors=True, views=False):
        """Streaming entry (zw_submit): starts the batch and returns a PendingBatch at once, or None when every
        pipeline slot of the context is taken (take `.result()` of an earlier batch first).  The batch must fit
        the device budget in one chunk."""
        self._check(params, color, container)
        arr, keep = self._as_images(images, color)
        ticket = C.c_int(-1)
        rc = self.lib.zw_submit(self.h, arr, len(images), int(params.lossy_quality), int(params.method), C.byref(ticket))
        if rc == 7:  # ZW_ERR_BUSY
            return None
        if rc != 0:
            _raise_for(rc, self.lib)
        return PendingBatch(self, ticket.value, (arr, keep), container, raise_errors, views)

    def stage(self, images, color=ColorType.Rgb8):
        arr, keep = self._as_images(images, color)
        rc = self.lib.zw_stage_batch(self.h, arr, len(images))
        if rc != 0:
            _raise_for(rc, self.lib)
        self._n = len(images)

    def encode_resident(self, params):
        t = _lib.ZwTiming()
        rc = self.lib.zw_encode_resident(self.h, int(params.lossy_quality), int(params.method), C.byref(t))
        if rc != 0:
            _raise_for(rc, self.lib)
        return t.as_dict()

    def download(self, container=True, raise_errors=True):
        outs = (_lib.ZwOutput * self._n)()
        t = _lib.ZwTiming()
        rc = self.lib.zw_download(self.h, outs, self._n, 1 if container else 0, C.byref(t))
        if rc != 0:
            self._collect(outs, False)
            _raise_for(rc, self.lib)
        return self._collect(outs, raise_errors), t.as_dict()

    def measure_int_peak(self):
        """Thread-level integer instructions per second of this GPU (issue-rate microbenchmark)."""
        x = C.c_double(0)
        rc = self.lib.zw_measure_int_peak(self.h, C.byref(x))
        if rc != 0:
            _raise_for(rc, self.lib)
        return x.value

    def dump_stage(self, index, name, dtype=np.uint8):
        n = C.c_size_t(0)
        rc = self.lib.zw_dump_stage(self.h, index, name.encode(), None, 0, C.byref(n))
        if rc not in (0, 4):
            _raise_for(rc, self.lib)
        buf = np.zeros(max(1, n.value), np.uint8)
        rc = self.lib.zw_dump_stage(self.h, index, name.encode(), buf.ctypes.data, buf.size, C.byref(n))
        if rc != 0:
            _raise_for(rc, self.lib)
        return buf[:n.value].view(dtype) if n.value else buf[:0].view(dtype)


class BatchPipeline:
    """Streaming batch entry on ONE context: `depth` batches in flight through zw_submit / zw_wait, so that the
    H2D copy of a batch runs under the kernels of the one before it and its D2H copy under the kernels of the one
    after it -- no extra contexts or host threads.  `submit` returns a PendingBatch (`.result()` -> (list of
    .webp bytes, timing dict)); when every slot is taken it first completes the oldest batch.  Results are
    bit-identical to `Context.encode_batch`."""

    def __init__(self, device=0, depth=3, views=False, **ctx_kwargs):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.depth = depth
        self.views = views  # hand the files back as memoryviews into one copy of the arena per batch
        self.ctx = Context(device, depth=depth, **ctx_kwargs)
        self._inflight = []

    def submit(self, images, params, color=ColorType.Rgb8, container=True):
        while True:
            p = self.ctx.submit(images, params, color, container, views=self.views)
            if p is not None:
                self._inflight.append(p)
                return p
            if not self._inflight:
                raise DeviceError(7, "no free pipeline slot")
            self._inflight.pop(0).result()  # completes (and caches) the oldest batch, freeing its slot

    def encode_batches(self, batches, params, color=ColorType.Rgb8, container=True):
        """Encode an iterable of batches; returns the list of per-batch outputs, in order."""
        futs = [self.submit(b, params, color, container) for b in batches]
        return [f.result()[0] for f in futs]

    def drain(self):
        """Complete every batch in flight, oldest first; returns their (outputs, timing) pairs."""
        res = [p.result() for p in self._inflight]
        self._inflight = []
        return res

    def close(self):
        for p in self._inflight:
            try:
                p.result()
            except Exception:
                pass
        self._inflight = []
        self.ctx.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class MultiContext:
    """One batch over several GPUs of one box (zw_multi_*): contiguous slices of ceil(n / G) images, one host
    thread + context per GPU inside the library, no collective; results come back in image order."""

    def __init__(self, devices, max_device_bytes=0, depth=0):
        self.lib = _lib.load()
        res = (C.c_int * 5)()
        res[0] = depth
        lim = _lib.ZwLimits(max_device_bytes, 0, res)
        devs = (C.c_int * len(devices))(*devices)
        self.h = self.lib.zw_multi_create(devs, len(devices), C.byref(lim))
        if not self.h:
            code = self.lib.zw_last_error()
            raise DeviceError(code, self.lib.zw_strerror(code).decode() + " (no CPU fallback exists)")
        self.devices = list(devices)

    def encode_batch(self, images, params, color=ColorType.Rgb8, container=True, raise_errors=True):
        Context._check(params, color, container)
        arr, keep = Context._as_images(images, color)
        outs = (_lib.ZwOutput * len(images))()
        tm = (_lib.ZwTiming * len(self.devices))()
        rc = self.lib.zw_multi_encode(self.h, arr, len(images), int(params.lossy_quality), int(params.method), 1 if container else 0, outs, tm)
        if rc != 0:
            Context._collect(self, outs, False)
            _raise_for(rc, self.lib)
        return Context._collect(self, outs, raise_errors), [t.as_dict() for t in tm]

    def close(self):
        if self.h:
            self.lib.zw_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}
_default_lock = threading.Lock()


def default_context(device=0):
    """One cached context per (thread, device): a zw_ctx is not thread-safe."""
    key = (threading.get_ident(), device)
    with _default_lock:
        if key not in _default_ctx:
            _default_ctx[key] = Context(device)
        return _default_ctx[key]


class WebPEncoder:
    """api.rs:1244-1398.  `writer` is a bytearray the encoder APPENDS to (like &mut Vec<u8>)."""

    def __init__(self, writer, device=0):
        self.writer = writer
        self.params = EncoderParams()
        self.device = device
        self.icc_profile = b""
        self.exif_metadata = b""
        self.xmp_metadata = b""

    def set_icc_profile(self, icc_profile):
        self.icc_profile = bytes(icc_profile)

    def set_exif_metadata(self, exif_metadata):
        self.exif_metadata = bytes(exif_metadata)

    def set_xmp_metadata(self, xmp_metadata):
        self.xmp_metadata = bytes(xmp_metadata)

    def set_params(self, params):
        self.params = params

    def encode(self, data, width, height, color):
        ctx = default_context(self.device)
        meta = None
        if self.icc_profile or self.exif_metadata or self.xmp_metadata:
            meta = {"icc": self.icc_profile, "exif": self.exif_metadata, "xmp": self.xmp_metadata}
        outs, _ = ctx.encode_batch([(bytes(data), width, height)], self.params, color, metadata=meta)
        self.writer += outs[0]


def _validate_buffer_size(size, width, height, bpp):  # api.rs:917-934: only "too small" is an error here
    expected = width * height * bpp
    if size < expected:
        raise InvalidBufferSize("buffer too small: got %d, expected %d" % (size, expected))


class EncoderConfig:
    """api.rs:488-672: dimension-independent, reusable configuration (builder).  Default: lossy, quality 75, method 4."""

    def __init__(self):
        self._quality = 75.0
        self._preset = Preset.Default
        self._lossless = False
        self._method = 4
        self._near_lossless = 100
        self._alpha_quality = 100
        self._exact = False
        self._target_size = 0
        self._use_sharp_yuv = False

    @classmethod
    def new(cls):
        return cls()

    @classmethod
    def new_lossless(cls):
        c = cls()
        c._lossless = True
        return c

    @classmethod
    def with_preset(cls, preset, quality):
        c = cls()
        c._preset, c._quality = preset, float(quality)  # not clamped here (api.rs:537-543)
        return c

    def quality(self, quality):
        self._quality = min(max(float(quality), 0.0), 100.0)
        return self

    def preset(self, preset):
        self._preset = preset
        return self

    def lossless(self, lossless):
        self._lossless = bool(lossless)
        return self

    def method(self, method):
        self._method = min(int(method), 6)
        return self

    def near_lossless(self, value):
        self._near_lossless = min(int(value), 100)
        return self

    def alpha_quality(self, quality):
        self._alpha_quality = min(int(quality), 100)
        return self

    def exact(self, exact):
        self._exact = bool(exact)
        return self

    def target_size(self, size):
        self._target_size = int(size)
        return self

    def sharp_yuv(self, enable):
        self._use_sharp_yuv = bool(enable)
        return self

    def get_quality(self):
        return self._quality

    def get_preset(self):
        return self._preset

    def is_lossless(self):
        return self._lossless

    def get_method(self):
        return self._method

    def to_params(self):
        """api.rs:633-640: predictor on, quality rounded like fast_math::roundf ((x + 0.5) as i32, then `as u8` saturates)."""
        q = min(max(int(np.float32(self._quality) + np.float32(0.5)), 0), 255)
        return EncoderParams(True, not self._lossless, q, self._method)

    def _encode(self, data, width, height, color, device=0, metadata=None):
        data = bytes(data)
        _validate_buffer_size(len(data), width, height, color.bytes_per_pixel())
        out = bytearray()
        enc = WebPEncoder(out, device)
        enc.set_params(self.to_params())
        if metadata:
            enc.set_icc_profile(metadata.get("icc", b""))
            enc.set_exif_metadata(metadata.get("exif", b""))
            enc.set_xmp_metadata(metadata.get("xmp", b""))
        enc.encode(data, width, height, color)
        return bytes(out)

    def encode_rgba(self, data, width, height, device=0):
        return self._encode(data, width, height, ColorType.Rgba8, device)

    def encode_rgb(self, data, width, height, device=0):
        return self._encode(data, width, height, ColorType.Rgb8, device)


class Encoder:
    """api.rs:703-914: `Encoder.new_rgba(data, w, h).quality(85).encode()`."""

    def __init__(self, data, width, height, color, device=0):
        self._data, self._width, self._height, self._color = data, width, height, color
        self._config = EncoderConfig()
        self._meta = {}
        self._device = device

    @classmethod
    def new_rgba(cls, data, width, height, device=0):
        return cls(data, width, height, ColorType.Rgba8, device)

    @classmethod
    def new_rgb(cls, data, width, height, device=0):
        return cls(data, width, height, ColorType.Rgb8, device)

    @classmethod
    def new_l8(cls, data, width, height, device=0):
        return cls(data, width, height, ColorType.L8, device)

    @classmethod
    def new_la8(cls, data, width, height, device=0):
        return cls(data, width, height, ColorType.La8, device)

    def config(self, config):
        self._config = config
        return self

    def icc_profile(self, profile):
        self._meta["icc"] = bytes(profile)
        return self

    def exif_metadata(self, data):
        self._meta["exif"] = bytes(data)
        return self

    def xmp_metadata(self, data):
        self._meta["xmp"] = bytes(data)
        return self

    def encode(self):
        return self._config._encode(self._data, self._width, self._height, self._color, self._device, self._meta)

    def encode_into(self, output):
        output += self.encode()

    def encode_to_writer(self, writer):
        writer.write(self.encode())


for _name in ("quality", "preset", "lossless", "method", "near_lossless", "alpha_quality", "exact", "target_size", "sharp_yuv"):
    def _fwd(self, value, _n=_name):
        getattr(self._config, _n)(value)
        return self
    setattr(Encoder, _name, _fwd)


def encode_batch(images, params, color=ColorType.Rgb8, device=0):
    """Convenience batch entry on the default context of `device`: list of images -> list of .webp bytes."""
    return default_context(device).encode_batch(images, params, color)[0]
