#!/usr/bin/env python3
"""Build exp_build/lib_chainprof.so: the library with clock64 phase counters in the pass-1 chroma chain (chroma_chain1 /
chroma_mb), for tools/gpu_chain_prof.py.  Patches a temporary copy of the sources; the tree is left untouched."""
import os, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "image_webp_b200", "csrc")
tmp = tempfile.mkdtemp()
work = os.path.join(tmp, "image_webp_b200", "csrc")
shutil.copytree(src, work, ignore=shutil.ignore_patterns("*.so"))
shutil.copytree(os.path.join(ROOT, "include"), os.path.join(tmp, "include"))
p = os.path.join(work, "zw_search.cuh")
s = open(p).read()
def sub(old, new):
    global s
    assert old in s, old[:50]
    s = s.replace(old, new, 1)
sub('__device__ ChromaOut chroma_mb(', '''__device__ unsigned long long g_chain_prof[16];
__device__ long long g_chain_last;
__device__ int g_chain_on;
#define CTICK(i) do { if (g_chain_on) { const long long _t = clock64(); if (lane == 0) { g_chain_prof[i] += (unsigned long long)(_t - g_chain_last); g_chain_last = _t; } } } while (0)
__device__ ChromaOut chroma_mb(''')
sub('  const int b8 = lane & 7,', '  CTICK(1);\n  const int b8 = lane & 7,')
sub('    int cost = (int)residual_cost_flat<2, 0>(q, 0, cc);\n', '    int cost = (int)residual_cost_flat<2, 0>(q, 0, cc);\n    CTICK(2);\n')
sub('    cost = red8_add(cost); sse = red8_add(sse); nzac = red8_add(nzac);', '    CTICK(3);\n    cost = red8_add(cost); sse = red8_add(sse); nzac = red8_add(nzac);')
sub('    // ---- transform_chroma_blocks: lane b (< 8) takes over block b of the winning mode ----', '    CTICK(4);')
sub('  if (lane < 8) W.dcbuf[lane] = c[0];', '  CTICK(5);\n  if (lane < 8) W.dcbuf[lane] = c[0];')
sub('  left_derr = new_left;\n  top_derr = new_top;', '  CTICK(6);\n  left_derr = new_left;\n  top_derr = new_top;')
sub('  R.uvnz = __ballot_sync(FULL, lane < 8 && W.nzflag[lane] != 0) & 0xffu;', '  CTICK(7);\n  R.uvnz = __ballot_sync(FULL, lane < 8 && W.nzflag[lane] != 0) & 0xffu;')
sub('  R.uv_mode = uv_mode;\n  __syncwarp();\n  return R;', '  R.uv_mode = uv_mode;\n  __syncwarp();\n  CTICK(8);\n  return R;')
sub('    const ChromaOut C = chroma_mb(W, SP, cc, mbx, mby, left_derr, top_derr, lane);\n    MbRecord* r = &P.rec1[gmb];', '    CTICK(0);\n    const ChromaOut C = chroma_mb(W, SP, cc, mbx, mby, left_derr, top_derr, lane);\n    MbRecord* r = &P.rec1[gmb];')
sub('    __syncwarp();\n    mbx = nx; mby = ny;\n  }\n}', '    __syncwarp();\n    CTICK(9);\n    mbx = nx; mby = ny;\n  }\n  __syncwarp();\n  if (lane == 0) g_chain_on = 0;\n  __syncwarp();\n}')
sub('  int mbx = 0, mby = 0;\n  for (u32 i = 0; i < nmb; i++) {\n    const u32 gmb = d.mb_off + i;\n    uint2 c_src = n_src;', '  int mbx = 0, mby = 0;\n  if (lane == 0) { g_chain_last = clock64(); g_chain_on = 1; }\n  __syncwarp();\n  for (u32 i = 0; i < nmb; i++) {\n    const u32 gmb = d.mb_off + i;\n    uint2 c_src = n_src;')
sub('    const SegParams& SP = seg4[seg_on ? c_seg : 0u];', '    const SegParams& SP = seg4[seg_on ? c_seg : 0u];\n    CTICK(10);')
sub('    __syncwarp();\n    if (lane < 8) {\n      W.uvws[(1 + lane) * 32] = (mbx == 0) ? 129 : W.left_u[1 + lane];', '    __syncwarp();\n    CTICK(11);\n    if (lane < 8) {\n      W.uvws[(1 + lane) * 32] = (mbx == 0) ? 129 : W.left_u[1 + lane];')
sub('    if (ahead && i + 1 < nmb) fetch(i + 1, nx, ny, n_src, n_top, n_derr, n_seg);', '    CTICK(12);\n    if (ahead && i + 1 < nmb) fetch(i + 1, nx, ny, n_src, n_top, n_derr, n_seg);')
open(p, "w").write(s)
with open(os.path.join(work, "zw_capi.cu"), "a") as f:
    f.write('''
extern "C" int zw_debug_chain_prof(unsigned long long* out, int reset) {
  unsigned long long z[16] = {0};
  if (reset) { cudaMemcpyToSymbol(zw::g_chain_prof, z, sizeof(z)); return 0; }
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, zw::g_chain_prof, sizeof(z));
  return 0;
}
''')
out = os.path.join(ROOT, "exp_build", "lib_chainprof.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off",
                       "-shared", "-o", out, "zw_capi.cu"], cwd=work)
shutil.rmtree(tmp)
print("built", out)
