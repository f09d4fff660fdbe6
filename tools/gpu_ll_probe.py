import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import image_webp_b200 as Z
import photo_inputs as PI
ctx = Z.Context(0)
for n in (1, 8, 128, 1024):
    imgs = list(PI.batch(n))
    os.environ["ZW_LL_SPLIT"] = "1"
    for _ in range(3):
        outs, t = ctx.encode_batch(imgs, Z.EncoderParams(), Z.ColorType.Rgb8)
    print(n, "huffman %.3f ms" % t["stats_ms"], "residual %.3f tokens %.3f bits %.3f emit %.3f" % (t["yuv_ms"], t["analysis_ms"], t["token_ms"], t["assemble_ms"]))
