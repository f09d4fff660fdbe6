#!/usr/bin/env python3
"""A/B of library builds under exp_build/ (tuning experiments): per-stage device times of the bench workload plus one
hash over all output files, so a variant that changes a byte is seen at once.
  python tools/gpu_variants.py            -> runs every exp_build/lib_*.so in a subprocess (ZW_LIB_PATH)
  python tools/gpu_variants.py --one      -> one measurement with the library ZW_LIB_PATH names"""
import glob, hashlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

def one():
    import image_webp_b200 as Z
    from image_webp_b200 import synth
    n = int(os.environ.get("ZW_VAR_N", "1024"))
    photo = os.environ.get("ZW_VAR_PHOTO", "0") == "1"
    if photo:
        import photo_inputs
        imgs = list(photo_inputs.batch(n))
    else:
        imgs = list(synth.batch_photo_like(n, 768, 512, 0))
    ctx = Z.Context(0)
    ctx.stage(imgs)
    p = Z.EncoderParams.lossy(75); p.method = 4
    ts = [ctx.encode_resident(p) for _ in range(4)]
    t = min(ts[1:], key=lambda x: x["device_total_ms"])
    outs, _ = ctx.download()
    h = hashlib.sha256()
    for o in outs: h.update(bytes(o))
    print("%-28s %s total %.2f  an %.2f p1 %.2f c1 %.2f st %.2f c2 %.2f p2 %.2f tok %.2f bc %.2f  sha %s" % (
        os.path.basename(os.environ.get("ZW_LIB_PATH", "default")), "photo" if photo else "synth", t["device_total_ms"], t["analysis_ms"], t["pass1_ms"], t["chroma1_ms"],
        t["stats_ms"], t["chroma2_ms"], t["pass2_ms"], t["token_ms"], t["boolcode_ms"], h.hexdigest()[:12]), flush=True)

if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        libs = [None] + sorted(glob.glob(os.path.join(ROOT, "exp_build", "lib_*.so")))
        for photo in ("0", "1") if "--photo" in sys.argv else ("0",):
            for lib in libs:
                env = dict(os.environ, ZW_VAR_PHOTO=photo)
                if lib: env["ZW_LIB_PATH"] = lib
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True, timeout=300)
                sys.stdout.write(r.stdout if r.returncode == 0 else "%s FAILED rc=%d %s\n" % (lib, r.returncode, r.stderr[-400:]))
                sys.stdout.flush()
