import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["ZW_LIB_PATH"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "exp_build", "lib_chainprof.so")
os.environ["ZW_NO_SIDE"] = "1"
import numpy as np
import image_webp_b200 as Z
from image_webp_b200 import synth, _lib
L = _lib.load()
ctx = Z.Context(0)
img = synth.photo_like(768, 512, 0)
ctx.stage([img])
p = Z.EncoderParams.lossy(75); p.method = 4
ctx.encode_resident(p)
buf = (C.c_ulonglong * 16)()
L.zw_debug_chain_prof(buf, 1)
t = ctx.encode_resident(p)
L.zw_debug_chain_prof(buf, 0)
names = ["fetch issue (after staging)", "dc (2 reduces)", "pred+fdct+quant+cost", "dequant+idct+sse", "3 reduces+score+select", "transfer shuffles",
         "dcbuf+diffusion", "final quant+idct+stores", "ballot", "record/border stores", "loop top (copies, SP)", "staging stores + sync 1",
         "left column + sync 2"]
tot = sum(buf[i] for i in range(13))
nmb = 48 * 32
print("chroma1 %.3f ms; %d cycles per MB" % (t["chroma1_ms"], tot / nmb))
for i, nm in enumerate(names):
    print("%-26s %6.0f cycles/MB  %4.1f%%" % (nm, buf[i] / nmb, 100.0 * buf[i] / tot))
