"""ctypes loader of the CUDA library (image_webp_b200/csrc/libzenwebp_b200.so) and its build recipe.

The library is built IN-TREE with nvcc for sm_100a; there is no CPU fallback: `load()` raises if
the shared object is missing and cannot be built, and every encode call raises if no CUDA device
is usable."""
import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libzenwebp_b200.so")
SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc")))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    hdr = os.path.join(os.path.dirname(HERE), "include", "zenwebp_b200.h")
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES) or os.path.getmtime(hdr) > t


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and %s is missing or stale: cannot build the CUDA library" % SO)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(CSRC, "zw_capi.cu")]
    subprocess.check_call(cmd, cwd=CSRC)
    return SO


class ZwImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("width", C.c_uint32), ("height", C.c_uint32),
                ("color", C.c_uint32), ("reserved", C.c_uint32)]


class ZwOutput(C.Structure):
    _fields_ = [("data", C.c_void_p), ("cap", C.c_size_t), ("len", C.c_size_t), ("status", C.c_int), ("reserved", C.c_int)]


class ZwLimits(C.Structure):
    _fields_ = [("max_device_bytes", C.c_size_t), ("persistent_warps_per_sm", C.c_int), ("reserved", C.c_int * 5)]


class ZwTiming(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "yuv_ms", "analysis_ms", "pass1_ms", "stats_ms", "pass2_ms", "token_ms",
                                         "boolcode_ms", "assemble_ms", "d2h_ms", "device_total_ms", "wall_ms")] + \
               [(n, C.c_uint64) for n in ("kernel_launches", "h2d_bytes", "d2h_bytes", "pixels")] + \
               [(n, C.c_float) for n in ("chroma1_ms", "chroma2_ms")] + [("symbols", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ZwBatchView(C.Structure):
    _fields_ = [("arena", C.c_void_p), ("n", C.c_size_t), ("offsets", C.POINTER(C.c_uint64)), ("lens", C.POINTER(C.c_uint32)),
                ("status", C.POINTER(C.c_int32))]


class ZwBlob(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t)]


class ZwDecodeInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32), ("filter_type", C.c_uint32),
                ("filter_level", C.c_uint32), ("sharpness", C.c_uint32), ("num_partitions", C.c_uint32),
                ("segments_enabled", C.c_uint32), ("sse_rgb", C.c_uint64), ("psnr_rgb", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ZwParams(C.Structure):
    _fields_ = [("use_predictor_transform", C.c_int), ("use_lossy", C.c_int), ("lossy_quality", C.c_int), ("method", C.c_int)]


class ZwMetadata(C.Structure):
    _fields_ = [("icc_profile", C.c_char_p), ("icc_len", C.c_size_t), ("exif", C.c_char_p), ("exif_len", C.c_size_t),
                ("xmp", C.c_char_p), ("xmp_len", C.c_size_t)]


# Every symbol include/zenwebp_b200.h declares.
EXPORTS = ["zw_create", "zw_destroy", "zw_last_error", "zw_strerror", "zw_free", "zw_max_output_size",
           "zw_encode_vp8_batch", "zw_encode_webp_batch", "zw_submit", "zw_wait", "zw_release",
           "zw_multi_create", "zw_multi_destroy", "zw_multi_device_count", "zw_multi_encode",
           "zw_stage_batch", "zw_encode_resident", "zw_download",
           "zw_dump_stage", "zw_version", "zw_measure_int_peak",
           "zw_decode_batch", "zw_verify", "zw_decode_dump_stage",
           "zw_encode_lossless_batch", "zw_encode_alpha_batch", "zw_params_default", "zw_encode_batch", "zw_lossless_dump_stage"]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("ZW_LIB_PATH")  # tuning experiments: load an alternative build
    if override:
        L = C.CDLL(override)
    else:
        if _stale():
            build()
        L = C.CDLL(SO)
    L.zw_create.restype = C.c_void_p
    L.zw_create.argtypes = [C.c_int, C.POINTER(ZwLimits)]
    L.zw_destroy.argtypes = [C.c_void_p]
    L.zw_strerror.restype = C.c_char_p
    L.zw_strerror.argtypes = [C.c_int]
    L.zw_version.restype = C.c_char_p
    L.zw_free.argtypes = [C.c_void_p]
    L.zw_max_output_size.restype = C.c_size_t
    L.zw_max_output_size.argtypes = [C.c_uint32, C.c_uint32]
    for f in ("zw_encode_vp8_batch", "zw_encode_webp_batch"):
        getattr(L, f).argtypes = [C.c_void_p, C.POINTER(ZwImage), C.c_size_t, C.c_int, C.c_int, C.POINTER(ZwOutput), C.POINTER(ZwTiming)]
    L.zw_submit.argtypes = [C.c_void_p, C.POINTER(ZwImage), C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.zw_wait.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ZwBatchView), C.POINTER(ZwTiming)]
    L.zw_release.argtypes = [C.c_void_p, C.c_int]
    L.zw_multi_create.restype = C.c_void_p
    L.zw_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(ZwLimits)]
    L.zw_multi_destroy.argtypes = [C.c_void_p]
    L.zw_multi_device_count.argtypes = [C.c_void_p]
    L.zw_multi_encode.argtypes = [C.c_void_p, C.POINTER(ZwImage), C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(ZwOutput), C.POINTER(ZwTiming)]
    L.zw_stage_batch.argtypes = [C.c_void_p, C.POINTER(ZwImage), C.c_size_t]
    L.zw_encode_resident.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ZwTiming)]
    L.zw_download.argtypes = [C.c_void_p, C.POINTER(ZwOutput), C.c_size_t, C.c_int, C.POINTER(ZwTiming)]
    L.zw_dump_stage.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.zw_measure_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.zw_decode_batch.argtypes = [C.c_void_p, C.POINTER(ZwBlob), C.c_size_t, C.c_int, C.POINTER(ZwOutput), C.POINTER(ZwImage),
                                  C.POINTER(ZwDecodeInfo), C.POINTER(C.c_float)]
    L.zw_verify.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ZwDecodeInfo), C.POINTER(C.c_float)]
    L.zw_decode_dump_stage.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.zw_encode_lossless_batch.argtypes = [C.c_void_p, C.POINTER(ZwImage), C.c_size_t, C.c_int, C.c_int, C.POINTER(ZwOutput), C.POINTER(ZwTiming)]
    L.zw_encode_alpha_batch.argtypes = [C.c_void_p, C.POINTER(ZwImage), C.c_size_t, C.POINTER(ZwOutput), C.POINTER(ZwTiming)]
    L.zw_params_default.restype = ZwParams
    L.zw_encode_batch.argtypes = [C.c_void_p, C.POINTER(ZwImage), C.c_size_t, C.POINTER(ZwParams), C.POINTER(ZwMetadata), C.POINTER(ZwOutput),
                                  C.POINTER(ZwTiming)]
    L.zw_lossless_dump_stage.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    _lib = L
    return L
