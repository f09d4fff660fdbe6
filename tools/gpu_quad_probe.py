#!/usr/bin/env python3
"""Quad vs warp luma kernels on the synthetic and photo workloads (stage times), with a byte comparison of the two."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import hashlib
    import image_webp_b200 as Z
    from image_webp_b200 import synth
    import photo_inputs as PI
    n = int(sys.argv[2])
    ctx = Z.Context(0)
    for name, imgs in (("synthetic", list(synth.batch_photo_like(n, 768, 512, 0))), ("photo", list(PI.batch(n)))):
        ctx.stage(imgs)
        for q, m in ((75, 4), (75, 6), (50, 0)):
            p = Z.EncoderParams.lossy(q); p.method = m
            for _ in range(2):
                t = ctx.encode_resident(p)
            outs, _ = ctx.download()
            h = hashlib.sha256(b"".join(hashlib.sha256(o).digest() for o in outs)).hexdigest()[:12]
            print("%s q%d m%d: total %.1f ms (%.0f MPix/s) pass1 %.1f pass2 %.1f  hash %s" % (name, q, m, t["device_total_ms"],
                  n * 768 * 512 / t["device_total_ms"] / 1e3, t["pass1_ms"], t["pass2_ms"], h), flush=True)
else:
    n = sys.argv[1] if len(sys.argv) > 1 else "1024"
    for mode in ("0", "1"):
        print("== ZW_QUAD=%s" % mode, flush=True)
        subprocess.run([sys.executable, __file__, "child", n], env=dict(os.environ, ZW_QUAD=mode))
