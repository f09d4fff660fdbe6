import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_webp_b200 as Z
from image_webp_b200 import synth
ctx = Z.Context(0); imgs = list(synth.batch_photo_like(1024, 768, 512, 0)); ctx.stage(imgs)
p = Z.EncoderParams.lossy(75); p.method = 0
ts = [ctx.encode_resident(p)["yuv_ms"] for _ in range(8)]
print("yuv_ms min %.4f -> %.0f GB/s" % (min(ts), 1024 * 768 * 512 * 4.5 / min(ts) / 1e6))
