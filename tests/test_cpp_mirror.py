"""include/zenwebp_b200.hpp (the C++ host-side mirror of the reference's encoder API) is compiled against the C ABI;
without a GPU its error path is checked, on a GPU one image goes through every class of it."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from image_webp_b200 import _lib, synth

ROOT = O.ROOT
SRC = os.path.join(ROOT, "tests", "cppcheck", "hpp_driver.cpp")
EXE = os.path.join(ROOT, "tests", "cppcheck", "_build", "hpp_driver")


def _build():
    so = _lib.build()
    deps = [SRC, os.path.join(ROOT, "include", "zenwebp_b200.hpp"), os.path.join(ROOT, "include", "zenwebp_b200.h"), so]
    if (not os.path.exists(EXE)) or any(os.path.getmtime(EXE) < os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", SRC, "-o", EXE, "-L" + os.path.dirname(so), "-lzenwebp_b200",
                               "-Wl,-rpath," + os.path.dirname(so), "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])
    return EXE


def test_hpp_compiles_and_fails_loudly_without_a_device():
    exe = _build()
    r = subprocess.run([exe, "selftest"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_hpp_every_class_encodes_byte_identically(tmp_path):
    exe = _build()
    img = synth.photo_like(160, 112, 77)
    raw = tmp_path / "in.rgb"
    raw.write_bytes(img.tobytes())
    r = subprocess.run([exe, "encode", "160", "112", str(raw), str(tmp_path / "out")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ref = O.encode(img, 75, 4)[1]
    for kind in ("batch", "encoder", "pipe", "multi"):
        assert (tmp_path / ("out.%s.webp" % kind)).read_bytes() == ref, kind
    assert (tmp_path / "out.lossless.webp").read_bytes() == O.webp_encode(img, "Rgb8")[1]
    assert (tmp_path / "out.meta.webp").read_bytes() == O.webp_encode(img, "Rgb8", use_lossy=True, quality=75, method=6, exif=b"EXI")[1]
    dec = np.frombuffer((tmp_path / "out.decoded.rgb").read_bytes(), np.uint8).reshape(112, 160, 3)
    assert np.array_equal(dec, O.decode(ref, True, ("rgb",))[1]["rgb"])
