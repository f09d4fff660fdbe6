#!/usr/bin/env python3
"""GPU bring-up helper: encode a list of cases through the C ABI, compare every stage with the
oracle, print the first mismatching stage per case and per-stage device times."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_lib as O  # noqa: E402
import parity_util as PU  # noqa: E402
import image_webp_b200 as Z  # noqa: E402
from image_webp_b200 import synth  # noqa: E402

CASES = [
    ("grad64 q75 m4", synth.gradient(64, 64), 75, 4),
    ("solid64 q75 m0", synth.solid(64, 64), 75, 0),
    ("photo64 q75 m0", synth.photo_like(64, 64, 1), 75, 0),
    ("photo64 q75 m2", synth.photo_like(64, 64, 1), 75, 2),
    ("photo64 q75 m4", synth.photo_like(64, 64, 1), 75, 4),
    ("odd99x87 q75 m4", synth.photo_like(99, 87, 2), 75, 4),
    ("photo256 q50 m0", synth.photo_like(256, 256, 1), 50, 0),
    ("photo320x272 q75 m4", synth.photo_like(320, 272, 21), 75, 4),
    ("photo768 q75 m4", synth.photo_like(768, 512, 0), 75, 4),
    ("photo768 q75 m6", synth.photo_like(768, 512, 0), 75, 6),
    ("noise256 q90 m4", synth.noise(256, 256, 3), 90, 4),
]


def main():
    sel = sys.argv[1:] or None
    ctx = Z.Context(0)
    nbad = 0
    for name, img, q, m in CASES:
        if sel and not any(s in name for s in sel):
            continue
        try:
            rc, ref, dump = O.encode(img, q, m, want_dump=True)
            p = Z.EncoderParams.lossy(q)
            p.method = m
            t0 = time.time()
            outs, t = ctx.encode_batch([img], p, raise_errors=False)
            dt = time.time() - t0
            gpu = outs[0]
            mbw = (img.shape[1] + 15) // 16
            rep = PU.compare_stages(ctx, 0, dump, mbw)
            same = gpu == ref
            print("[%s] %s  gpu %s B oracle %d B  wall %.1f ms  dev %.2f ms (yuv %.3f an %.3f p1 %.2f st %.2f p2 %.2f tok %.2f bc %.2f)" %
                  ("OK " if same and not rep else "BAD", name, len(gpu) if gpu else None, len(ref), dt * 1e3, t["device_total_ms"],
                   t["yuv_ms"], t["analysis_ms"], t["pass1_ms"], t["stats_ms"], t["pass2_ms"], t["token_ms"], t["boolcode_ms"]))
            if rep or not same:
                nbad += 1
                for r in rep[:4]:
                    print("    " + r.replace("\n", "\n    "))
        except Exception:
            nbad += 1
            print("[EXC] %s" % name)
            traceback.print_exc()
            break
        sys.stdout.flush()
    print("mismatching cases: %d" % nbad)
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
