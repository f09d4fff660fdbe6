#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_golden.json and the two small .webp fixtures from the CPU
oracle.  The reference (Rust) cannot be run in this image, so these vectors pin the ORACLE
against drift (any edit that changes its bytes fails the CPU suite); they are not outputs of the
reference binary.  Run:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import photo_inputs  # noqa: E402
from image_webp_b200 import synth  # noqa: E402

CASES = [
    ("grad64_q75_m4", "gradient", (64, 64), 75, 4),
    ("chk128_q50_m4", "checker_gradient", (128, 128), 50, 4),
    ("chk128_q90_m4", "checker_gradient", (128, 128), 90, 4),
    ("solid64_q75_m4", "solid", (64, 64), 75, 4),
    ("noise64_q75_m4", "noise", (64, 64), 75, 4),
    ("photo99x87_s2_q75_m4", "photo_like", (99, 87, 2), 75, 4),
    ("photo256_s1_q50_m0", "photo_like", (256, 256, 1), 50, 0),
    ("photo256_s7_q75_m2", "photo_like", (256, 256, 7), 75, 2),
    ("photo256_s8_q75_m5", "photo_like", (256, 256, 8), 75, 5),
    ("photo320x272_s21_q75_m4", "photo_like", (320, 272, 21), 75, 4),
    ("photo768_s0_q75_m4", "photo_like", (768, 512, 0), 75, 4),
    ("photo768_s0_q75_m6", "photo_like", (768, 512, 0), 75, 6),
    ("photo768_s9_q90_m4", "photo_like", (768, 512, 9), 90, 4),
    ("photo17_s4_q50_m2", "photo_like", (17, 17, 4), 50, 2),
    # real photographs (tests/golden/photos, the reference's gallery1 PNGs): SURVEY.md 8(d) config 1(i) and friends
    ("real3_256_104_q75_m4", "crop", ("3", 256, 104), 75, 4),
    ("real3_256_104_q75_m6", "crop", ("3", 256, 104), 75, 6),
    ("real4_100_200_q90_m4", "crop", ("4", 100, 200), 90, 4),
    ("real5_0_0_q50_m0", "crop", ("5", 0, 0), 50, 0),
    ("real5_200_240_q75_m1", "crop", ("5", 200, 240), 75, 1),
    ("real4_full_q75_m4", "photo", ("4",), 75, 4),
]


def build_image(kind, args):
    return (getattr(synth, kind, None) or getattr(photo_inputs, kind))(*args)


def main():
    out = {}
    for name, kind, args, q, m in CASES:
        img = build_image(kind, args)
        rc, data, dump = O.encode(img, q, m, want_dump=True)
        assert rc == 0
        out[name] = {"kind": kind, "args": list(args), "quality": q, "method": m, "bytes": len(data),
                     "sha256": hashlib.sha256(data).hexdigest(),
                     "input_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
                     "part0_bytes": int(dump["PART0"].size), "part1_bytes": int(dump["PART1"].size),
                     "i4_mbs": int((dump["P2MB"]["ymode"] == 4).sum()), "skipped_mbs": int(dump["P2MB"]["skip"].sum())}
        if name in ("grad64_q75_m4", "photo99x87_s2_q75_m4"):
            open(os.path.join(HERE, name + ".webp"), "wb").write(data)
    json.dump(out, open(os.path.join(HERE, "oracle_golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote %d cases" % len(out))


if __name__ == "__main__":
    main()
