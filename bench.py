#!/usr/bin/env python3
"""bench.py -- lossy encode MPix/s (q75 m4, byte-identical to the CPU port) on N B200s, beside the CPU baseline.

Workload (BASELINE.json configs[1]): a batch of 1024 x 768x512 RGB images, quality 75, method 4,
synthetic "photo-like" content (image_webp_b200/synth.py).  One step = one encode of the whole
batch.  Images are independent, so N GPUs each encode their own batch of 1024 (weak scaling, no
collective on the data path); the metric is Σ pixels of all ranks / max-over-ranks time.

  value : kernel-only, inputs already resident in HBM (zw_encode_resident), CUDA-event time
  e2e   : the streaming entry of the C ABI (zw_submit / zw_wait / zw_release on ONE context, via
          BatchPipeline): pinned HOST RGB in, .webp bytes out; every step's H2D and D2H copies
          are inside the timed region (they overlap the neighbouring steps' kernels)
  photo : the same two figures on 1024 DISTINCT 768x512 crops of the reference's own test
          photographs (tests/golden/photos) -- real photographs carry 2-5x the token symbols of the
          synthetic generator; EVERY output of that batch is compared with the multi-threaded oracle
  other_configs : BASELINE.json configs 1, 3, 4, 5 (kernel-only, end to end, parity sample, CPU port)
  --impl reference : the CPU oracle (a C++ restatement of the reference; the Rust reference cannot
                     be built in this image) on all host cores, one image per thread.
"byte-identical" everywhere means: identical to the CPU port under oracle/ (the Rust reference cannot
be run here)."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, QUALITY, METHOD = 768, 512, 75, 4
WORKLOAD = "batch of %d x 768x512 RGB, q75 method 4"
METRIC = "lossy encode MPix/s (q75 m4, byte-identical)"
# algorithmic HBM bytes per pixel of the dominant kernel (pass-2 luma search k_search<2>): read the Y
# plane (1 B/px) + write the 32-byte header and 17 luma blocks (544 B) of one macroblock record per
# 256 px (2.25 B/px).  DESIGN.md §Measurement.
SEARCH_BYTES_PER_PX = 1.0 + 576.0 / 256.0
# the reference's own published single-thread figure for this exact metric (768x512 Kodak photo, q75 m4, SIMD Rust
# build, hardware unstated): /root/reference/CLAUDE.md:14, BASELINE.md section 1
REF_PUBLISHED_MPIX_S_PER_THREAD = 6.2
STAGES = ("yuv_ms", "analysis_ms", "pass1_ms", "chroma1_ms", "stats_ms", "chroma2_ms", "pass2_ms", "token_ms", "boolcode_ms", "assemble_ms")


def rank_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def native_oracle():
    """The oracle rebuilt with -O3 -march=native on THIS box for the CPU baseline (falls back to the
    generic build that travels with the repo)."""
    import ctypes as C
    import oracle_lib as O
    so = os.path.join(ROOT, "oracle", "_build", "libzw_oracle_native.so")
    src = os.path.join(ROOT, "oracle", "zw_oracle.cpp")
    try:
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread", "-shared",
                                   "-o", so, src], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = C.CDLL(so)
    except Exception:
        L = O.lib()
    L.zwo_encode_batch_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.zwo_encode_batch_mt.restype = C.c_size_t
    L.zwo_encode_batch_mt_out.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_size_t, C.c_void_p]
    L.zwo_encode_batch_mt_out.restype = C.c_size_t
    return L


def cpu_run(imgs, threads, lib, quality=QUALITY, method=METHOD):
    """Encode imgs ([n,h,w,3] uint8) with the oracle, one image per thread at a time.  Returns seconds."""
    n, h, w, _ = imgs.shape
    t0 = time.perf_counter()
    lib.zwo_encode_batch_mt(imgs.ctypes.data, n, w, h, quality, method, threads)
    return time.perf_counter() - t0


def cpu_baseline(sample, lib, cores, what, quality=QUALITY, method=METHOD):
    import numpy as np
    sample = np.ascontiguousarray(sample)
    n, h, w = sample.shape[:3]
    cpu_run(sample[:min(n, cores)], cores, lib, quality, method)
    ts = cpu_run(sample, cores, lib, quality, method)
    k = min(n, 4)
    t1 = cpu_run(sample[:k], 1, lib, quality, method)
    return {"value": n * w * h / ts / 1e6, "unit": "MPix/s", "cores": cores, "kind": "port",
            "single_thread_mpix_s": k * w * h / t1 / 1e6,
            "reference_published_single_thread_mpix_s": REF_PUBLISHED_MPIX_S_PER_THREAD,
            "note": "kind=port: the scalar C++ restatement of the reference (oracle/, -O3 -march=native); the reference's own SIMD Rust "
                    "build publishes %.1f MPix/s per thread for q75 m4 on a 768x512 photograph (reference CLAUDE.md:14, hardware "
                    "unstated) and cannot be built here (no Rust toolchain)" % REF_PUBLISHED_MPIX_S_PER_THREAD,
            "sample": "%d %s, one image per thread on %d threads" % (n, what, cores)}


def run_reference(args):
    rank, _, world = rank_env()
    if rank != 0:
        return 0
    from image_webp_b200 import synth
    cores = os.cpu_count() or 1
    lib = native_oracle()
    n_sample = max(cores * 16, 128)  # ~2-4 s of all-core CPU work per step
    imgs = synth.batch_photo_like(n_sample, W, H, 0)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_run(imgs[:cores], cores, lib)
    times = [cpu_run(imgs, cores, lib) for _ in range(args.steps)]
    tsum = sum(times)
    value = n_sample * W * H * args.steps / tsum / 1e6
    line = {"metric": METRIC, "value": value, "unit": "MPix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tsum / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD % 1024, "quality": QUALITY, "method": METHOD,
                       "note": "bounded sample of the workload per step; CPU only"},
            "cpu_baseline": {"value": value, "unit": "MPix/s", "cores": cores, "kind": "port",
                             "reference_published_single_thread_mpix_s": REF_PUBLISHED_MPIX_S_PER_THREAD,
                             "sample": "%d of the 1024 768x512 images per step, one image per thread on %d threads" % (n_sample, cores)},
            "e2e": {"value": value, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def pinned_batch(torch, arr):
    host = torch.empty(arr.shape, dtype=torch.uint8, pin_memory=True)
    host.numpy()[...] = arr
    return host


def run_workload(Z, torch, ctx, pipe, imgs, params, steps, warmup, barrier, sampler=None):
    """Kernel-only (resident) and end-to-end (streaming C ABI) legs of one workload.  Returns a dict of raw sums."""
    ctx.stage(imgs)
    for _ in range(warmup):
        ctx.encode_resident(params)
    barrier()
    if sampler:
        sampler.start()
    stage_ms = {k: 0.0 for k in STAGES}
    dev_ms, launches, symbols = 0.0, 0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        t = ctx.encode_resident(params)
        dev_ms += t["device_total_ms"]
        launches += t["kernel_launches"]
        symbols = t["symbols"]
        for k in STAGES:
            stage_ms[k] += t[k]
    barrier()
    wall_kernel_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    outs_resident, _ = ctx.download()
    # end to end: (a) the streaming entry, `depth` batches in flight on one context.  The zw_image descriptors of the
    # (unchanging) pinned input buffers are built once; every step still copies all pixels host -> device and all files
    # device -> host (+ one host copy out of the pinned arena) inside the timed region.
    prep = pipe.ctx.prepare(imgs)
    for f in [pipe.submit(prep, params) for _ in range(pipe.depth)]:
        f.result()
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_dev_ms = 0.0
    outs = None
    futs = []
    for _ in range(steps):
        futs.append(pipe.submit(prep, params))
    for f in futs:
        outs, t = f.result()
        h2d += t["h2d_bytes"]; d2h += t["d2h_bytes"]; e2e_dev_ms += t["device_total_ms"]
    barrier()
    e2e_s = time.perf_counter() - t0
    outs = [bytes(o) for o in outs]
    # (b) one blocking call per step (zw_encode_webp_batch: chunks pipelined inside the call)
    pipe.ctx.encode_batch(imgs, params)
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, steps // 2)):
        outs_b, tb = pipe.ctx.encode_batch(imgs, params)
    barrier()
    e2e_call_s = (time.perf_counter() - t0) / max(1, steps // 2) * steps
    return {"dev_s": dev_ms / 1e3, "wall_kernel_s": wall_kernel_s, "e2e_s": e2e_s, "e2e_call_s": e2e_call_s, "stage_ms": stage_ms,
            "launches": launches, "symbols": symbols, "h2d": h2d, "d2h": d2h, "clocks": clocks, "e2e_dev_ms": e2e_dev_ms,
            "outs": outs, "outs_resident": outs_resident, "outs_call": outs_b, "h2d_ms": t["h2d_ms"], "d2h_ms": t["d2h_ms"]}


def verify_leg(Z, ctx, workloads, params, check=4):
    """The on-device VP8 decoder as batch verifier (zw_verify): every file of a freshly encoded batch is decoded where it
    lies in device memory and scored against the source pixels resident there; `check` images per workload are also decoded
    by the decoder oracle (CPU restatement of the reference's decoder) and must give the same squared error."""
    import numpy as np
    import oracle_lib as O
    res = {"api": "zw_submit -> zw_verify(ticket) -> zw_wait: parse (one warp per image), reconstruct + loop filter (row wavefronts), "
                  "colour conversion + squared error; nothing but the per-image results crosses the link",
           "decoder_pinned_by": "the reference's pixel-exact decode fixtures (tests/decode.rs gallery1, gallery1_nofancy, gallery2 *_webp_a) via "
                                "the decoder oracle; GPU == oracle at every stage (tests/test_gpu_decoder.py)"}
    for name, imgs in workloads:
        n = len(imgs)
        px = sum(i.shape[0] * i.shape[1] for i in imgs)
        best = None
        for _ in range(3):
            pend = ctx.submit(imgs, params)
            info, ms = Z.verify_pending(pend)
            outs, t = pend.result()
            if best is None or sum(ms) < sum(best):
                best = ms
        ps = np.array([i["psnr_rgb"] for i in info])
        ok = sum(1 for i in info if i["status"] == 0)
        same = 0
        idx = list(range(0, n, max(1, n // check)))[:check]
        for i in idx:
            rc, o = O.decode(outs[i], True, ("rgb",))
            same += int(rc == 0 and info[i]["sse_rgb"] == int(((o["rgb"].astype(np.int64) - imgs[i].astype(np.int64)) ** 2).sum()))
        res[name] = {"images": n, "decode_mpix_s": px / sum(best) / 1e3, "ms": {"parse": best[0], "reconstruct": best[1], "filter": best[2], "colour": best[3]},
                     "encode_ms": t["device_total_ms"], "decoded_ok": ok, "psnr_rgb_min": float(ps.min()), "psnr_rgb_mean": float(ps.mean()),
                     "sse_equal_to_oracle_decode": {"checked": len(idx), "identical": same}}
        assert ok == n and same == len(idx), "verification leg: device decoder disagrees with the oracle"
    return res


def lossless_leg(Z, ctx, workloads, lib, cores, check=64):
    """The lossless (VP8L) path of WebPEncoder::encode on the same 1024-image workloads: device time of the
    seven kernels, end to end through zw_encode_batch from pinned host memory (H2D + D2H inside), a sample of the files
    checked byte for byte against the lossless oracle, and the oracle timed on all host cores."""
    import numpy as np
    import oracle_lib as O
    res = {"api": "zw_encode_batch (EncoderParams::default(): lossless, predictor transform on), one blocking call per batch through the raw C ABI "
                  "with caller-provided output buffers; four chunks pipelined over two slots inside the call",
           "pinned_by": "libwebp decodes the oracle's files to exactly the input pixels (the reference's own acceptance test, "
                        "api.rs:1447-1511; tests/test_oracle_lossless.py); GPU == oracle byte for byte (tests/test_gpu_lossless.py)"}
    import ctypes as C
    from image_webp_b200 import _lib
    L = _lib.load()
    zp = L.zw_params_default()
    for name, imgs in workloads:
        n = len(imgs)
        px = sum(int(im.shape[0]) * int(im.shape[1]) for im in imgs)
        prep = ctx.prepare(imgs)
        cap = max(int(im.shape[0]) * int(im.shape[1]) for im in imgs) * 5 + 4096  # far above any real file (noise stays below 4.2 B/px)
        store = np.empty((n, cap), np.uint8)   # caller-provided output buffers, reused by every call
        zouts = (_lib.ZwOutput * n)()
        best, best_wall = None, None
        for _ in range(3):
            for i in range(n):
                zouts[i].data = store[i].ctypes.data; zouts[i].cap = cap; zouts[i].len = 0
            t = _lib.ZwTiming()
            t0 = time.perf_counter()
            rc = L.zw_encode_batch(ctx.h, prep.arr, n, C.byref(zp), None, zouts, C.byref(t))
            wall = time.perf_counter() - t0
            assert rc == 0 and all(zouts[i].status == 0 for i in range(n)), "lossless %s: call failed" % name
            t = t.as_dict()
            if best is None or t["device_total_ms"] < best["device_total_ms"]:
                best = t
            best_wall = wall if best_wall is None else min(best_wall, wall)
        outs = [store[i, :zouts[i].len].tobytes() for i in range(n)]
        # kernel-only: the whole batch as ONE chunk (the pipelined call pays the Huffman kernel's latency once per chunk,
        # hidden behind the other slot's copies there but not in a sum of per-chunk device times)
        os.environ["ZW_LL_SPLIT"] = "1"
        try:
            for _ in range(2):
                t1 = _lib.ZwTiming()
                rc = L.zw_encode_batch(ctx.h, prep.arr, n, C.byref(zp), None, zouts, C.byref(t1))
                assert rc == 0
            best = t1.as_dict()
        finally:
            del os.environ["ZW_LL_SPLIT"]
        k = min(check, n)
        sample = np.stack([np.asarray(im) for im in imgs[:k]])
        ref, dt = O.webp_encode_batch_mt(sample, threads=cores, L=lib)
        bad = [i for i in range(k) if outs[i] != ref[i]]
        traffic = 3 + 4 + 4 + 2 + 6 + 6  # B/px: source read, residual written, read 3x, descriptors written, read 2x (+ the files)
        out_bytes = sum(len(o) for o in outs)
        res[name] = {"images": n, "kernel_only_mpix_s": px / best["device_total_ms"] / 1e3, "kernel_ms": best["device_total_ms"],
                     "e2e_mpix_s": px / best_wall / 1e6, "e2e_ms": 1e3 * best_wall,
                     "stage_ms": {"transforms": best["yuv_ms"], "tokens_hist": best["analysis_ms"], "huffman": best["stats_ms"],
                                  "bit_counts": best["token_ms"], "pack": best["assemble_ms"], "h2d": best["h2d_ms"], "d2h": best["d2h_ms"]},
                     "bytes_per_px": out_bytes / px, "h2d_bytes": best["h2d_bytes"], "d2h_bytes": best["d2h_bytes"],
                     "hbm_gb_s": (px * traffic + 2 * out_bytes) / best["device_total_ms"] / 1e6,
                     "hbm_frac_of_measured_peak": (px * traffic + 2 * out_bytes) / best["device_total_ms"] / 1e6 / measured_peaks()[0]["hbm_gbs"],
                     "parity": {"checked": k, "identical": k - len(bad), "against": "lossless oracle (oracle/zw_lossless_oracle.inc)"},
                     "cpu_port": {"value": k * px / n / dt / 1e6, "unit": "MPix/s", "cores": cores, "kind": "port",
                                  "sample": "%d images, one image per thread on %d threads" % (k, cores)}}
        assert not bad, "lossless %s: images %s differ from the oracle" % (name, bad[:8])
    return res


def other_configs(Z, torch, ctx, pipe, lib, cores, world):
    """BASELINE.json configs 1, 3, 4, 5 on one GPU (kernel-only + end to end + parity sample + CPU port)."""
    import numpy as np
    import oracle_lib as O
    import photo_inputs as PI
    from image_webp_b200 import synth
    res = {}

    def run(name, imgs, q, m, check, cpu_sample, reps=3):
        p = Z.EncoderParams.lossy(q)
        p.method = m
        px = sum(i.shape[0] * i.shape[1] for i in imgs)
        ctx.stage(imgs)
        ctx.encode_resident(p)
        ts = [ctx.encode_resident(p) for _ in range(reps)]
        dev = min(t["device_total_ms"] for t in ts)
        tb = ts[-1]
        prep = pipe.ctx.prepare(imgs)  # descriptors built once: the end-to-end leg times the blocking batch call itself
        pipe.ctx.encode_batch(prep, p)
        torch.cuda.synchronize()
        e2e = None
        for _ in range(reps):
            t0 = time.perf_counter()
            zouts, rc = pipe.ctx.encode_batch_raw(prep, p)
            dt = time.perf_counter() - t0
            assert rc == 0
            e2e = dt if e2e is None else min(e2e, dt)
            outs = pipe.ctx._collect(zouts, True)
        ok = 0
        for i in check:
            rc, ref, _ = O.encode(imgs[i], q, m)
            ok += int(rc == 0 and outs[i] == ref)
        r = {"pixels": px, "kernel_only_ms": dev, "kernel_only_mpix_s": px / dev / 1e3, "e2e_ms": 1e3 * e2e, "e2e_mpix_s": px / e2e / 1e6,
             "stage_ms": {k: tb[k] for k in STAGES}, "symbols_per_px": tb["symbols"] / px,
             "parity": {"checked": len(check), "identical": ok}}
        if cpu_sample is not None:
            cb = cpu_baseline(cpu_sample, lib, cores, "images of this config", q, m)
            r["cpu_port"] = {k: cb[k] for k in ("value", "cores", "single_thread_mpix_s", "sample")}
        res[name] = r
        assert ok == len(check), "%s: GPU output differs from the oracle" % name

    def pinned(arrs):  # inputs of the end-to-end legs live in pinned host memory, like the main workload's
        return pinned_batch(torch, np.stack(arrs)).numpy()

    c1 = pinned([PI.survey_crop()])
    run("config1_photo: 1 x 768x512 photograph crop (gallery1/3.png @256,104) q75 m4", [c1[0]], 75, 4, [0], np.stack([c1[0]] * 4), reps=5)
    s1 = pinned([synth.photo_like(768, 512, 0)])
    run("config1_synthetic: 1 x 768x512 G(0) q75 m4", [s1[0]], 75, 4, [0], np.stack([s1[0]] * 4), reps=5)
    big = pinned([synth.photo_like(4096, 4096, 3, freq_scale=4.0)])
    run("config3: 1 x 4096x4096 q90 m4", [big[0]], 90, 4, [0], big, reps=2)
    del big
    base = pinned([synth.photo_like(1920, 1080, 100 + i) for i in range(8)])
    run("config4: 256 x 1920x1080 q75 m6 (8 distinct images, 32 times each)", [base[i % 8] for i in range(256)], 75, 6, [0, 255],
        base[:max(2, min(8, cores // 2))], reps=2)
    tb_ = pinned([synth.photo_like(256, 256, 200 + i) for i in range(64)])
    run("config5_shard: 8192 x 256x256 q50 m0 (the 1/8 shard one GPU owns; 64 distinct images)", [tb_[i % 64] for i in range(8192)], 50, 0,
        [0, 63, 8191], tb_, reps=2)
    return res


def config5_multi(Z, torch, n_dev):
    """Config 5 through the library's own multi-GPU entry (zw_multi_encode) in ONE process: 8192 thumbnails per GPU
    (65 536 at 8 GPUs), sharded by image, gathered in image order."""
    import numpy as np
    import oracle_lib as O
    from image_webp_b200 import synth
    n = 8192 * n_dev
    block = pinned_batch(torch, np.stack([synth.photo_like(256, 256, 200 + i) for i in range(64)]))
    imgs = [block.numpy()[i % 64] for i in range(n)]
    p = Z.EncoderParams.lossy(50)
    p.method = 0
    import ctypes as C
    from image_webp_b200 import _lib
    mc = Z.MultiContext(list(range(n_dev)))
    prep = Z.Context.prepare(imgs)        # zw_image descriptors built once: the timed region is the C call alone
    try:
        mc.encode_batch(prep, p)
        zouts = (_lib.ZwOutput * n)()     # data == NULL: the library allocates every file (malloc) and fills it
        tm = (_lib.ZwTiming * n_dev)()
        t0 = time.perf_counter()
        rc = mc.lib.zw_multi_encode(mc.h, prep.arr, n, 50, 0, 1, zouts, tm)
        dt = time.perf_counter() - t0
        assert rc == 0
        outs = [C.string_at(zouts[i].data, zouts[i].len) if zouts[i].status == 0 else None for i in range(n)]
        for i in range(n):
            mc.lib.zw_free(zouts[i].data)
        tm = [t.as_dict() for t in tm]
    finally:
        mc.close()
    ok = sum(int(outs[i] == O.encode(imgs[i], 50, 0)[1]) for i in (0, 63, n // 2 + 5, n - 1))
    assert ok == 4 and all(o is not None for o in outs)
    px = n * 256 * 256
    return {"api": "zw_multi_encode (one process, one host thread + context per GPU, no collective)", "n_gpus": n_dev, "images": n,
            "e2e_ms": 1e3 * dt, "e2e_mpix_s": px / dt / 1e6, "per_gpu_kernel_ms": [t["device_total_ms"] for t in tm],
            "parity": {"checked": 4, "identical": ok}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-photo", action="store_true", help="skip the photo workload")
    ap.add_argument("--no-other", action="store_true", help="skip other_configs")
    ap.add_argument("--no-verify", action="store_true", help="skip the on-device decode / verification leg")
    ap.add_argument("--depth", type=int, default=3, help="batches in flight in the end-to-end leg (pipeline slots of the context)")
    ap.add_argument("--check", type=int, default=16, help="synthetic images per step byte-compared with the oracle after timing")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    rank, local_rank, world = rank_env()
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        # the ranks that finish early wait for rank 0's extra legs on the CPU (gloo): an NCCL barrier would park a
        # spinning kernel on their GPUs, which rank 0's one-process multi-GPU leg (config5_multi) is about to use
        tail_group = dist.new_group(backend="gloo")
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    import image_webp_b200 as Z
    from image_webp_b200 import synth
    n = args.batch
    params = Z.EncoderParams.lossy(QUALITY)
    params.method = METHOD
    ctx = Z.Context(dev)                      # kernel-only leg (split API, lane 0)
    # end-to-end leg: ONE context, `depth` batches in flight; the files of a batch are handed back as views into one
    # host copy of the slot's pinned arena (one memcpy per batch, no per-image Python objects with their own copies)
    pipe = Z.BatchPipeline(dev, depth=args.depth, views=True)
    pix = n * W * H
    cores = os.cpu_count() or 1

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic workload (the BASELINE config): distinct content per rank, pinned host staging ------------
    host = pinned_batch(torch, synth.batch_photo_like(n, W, H, seed0=1000 * rank))
    imgs = [host.numpy()[i] for i in range(n)]
    R = run_workload(Z, torch, ctx, pipe, imgs, params, args.steps, args.warmup, barrier, ClockSampler(dev))
    int_peak = ctx.measure_int_peak()  # thread-level integer instructions / s (microbenchmark, outside the timed regions)

    parity = None
    if rank == 0 and args.check > 0:
        import oracle_lib as O
        idx = list(range(0, n, max(1, n // args.check)))[:args.check]
        ref = O.encode_batch_mt(host.numpy()[idx], QUALITY, METHOD)
        ok = sum(int(R["outs"][i] == r and R["outs_resident"][i] == r and R["outs_call"][i] == r) for i, r in zip(idx, ref))
        parity = {"checked": len(idx), "identical": ok, "against": "CPU port (oracle/), three GPU entry points each"}
        assert ok == len(idx), "GPU output differs from the oracle"
    bytes_per_px = sum(len(o) for o in R["outs"]) / pix

    # ---- photo workload: 1024 distinct crops of real photographs, every output checked ---------------------------
    P = None
    photo_host = None
    if not args.no_photo:
        import photo_inputs as PI
        photo_host = torch.empty((n, H, W, 3), dtype=torch.uint8, pin_memory=True)
        PI.batch(n, W, H, out=photo_host.numpy(), first=(n * rank) % 1024)
        pimgs = [photo_host.numpy()[i] for i in range(n)]
        P = run_workload(Z, torch, ctx, pipe, pimgs, params, args.steps, args.warmup, barrier)

    # ---- reduce over ranks: max time -------------------------------------------------------------------------
    tl = [R["dev_s"], R["e2e_s"], R["wall_kernel_s"], R["e2e_call_s"]] + ([P["dev_s"], P["e2e_s"], P["e2e_call_s"]] if P else [0, 0, 0])
    times = torch.tensor(tl, dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, e2e_s, wall_kernel_s, e2e_call_s, p_dev_s, p_e2e_s, p_call_s = [float(x) for x in times.tolist()]
    total_pix = pix * world * args.steps
    value = total_pix / dev_s / 1e6
    e2e_value = total_pix / e2e_s / 1e6

    if rank == 0:
        peaks, how = measured_peaks()
        stage_ms = R["stage_ms"]
        dom = max(STAGES, key=lambda k: stage_ms[k])
        p2_s = stage_ms["pass2_ms"] / args.steps / 1e3   # k_search<2> alone (CUDA events around that launch)
        p1_s = stage_ms["pass1_ms"] / args.steps / 1e3
        yuv_s = stage_ms["yuv_ms"] / args.steps / 1e3
        prof = {}
        try:  # figures of the committed ncu --set full capture of this round (tools/ncu_summary.py)
            prof = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        except Exception:
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            except Exception:
                prof = {}

        def prof_of(sub):
            for kname, v in prof.get("kernels", {}).items():
                if sub in kname:
                    return v
            return {}
        ps2, pyuv = prof_of("k_search<(int)2>") or prof_of("k_search<2>"), prof_of("k_yuv")
        traffic = ps2.get("dram_bytes_per_pixel", 0) * pix or None
        traffic_yuv = pyuv.get("dram_bytes_per_pixel", 0) * pix or None
        line = {
            "metric": METRIC, "value": value, "unit": "MPix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": WORKLOAD % n, "images_per_gpu": n, "quality": QUALITY, "method": METHOD,
                       "l2": "inputs (%.2f GB per step) larger than L2" % (n * W * H * 3 / 1e9), "timing": "cuda events on the library stream, max over ranks",
                       "byte_identical_to": "the CPU port of the reference (oracle/); the Rust reference itself cannot be run here"},
            "e2e": {"value": e2e_value, "unit": "MPix/s", "h2d_bytes_per_step": R["h2d"] // args.steps, "d2h_bytes_per_step": R["d2h"] // args.steps,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "frac_of_kernel_only": e2e_value / value,
                    "kernel_ms_per_step_while_streaming": R["e2e_dev_ms"] / args.steps,
                    "api": "zw_submit / zw_wait / zw_release on one context, %d batches in flight (BatchPipeline): pinned host RGB in, .webp bytes out" % args.depth,
                    "h2d_gb_s": R["h2d"] / args.steps / max(R["h2d_ms"], 1e-6) / 1e6,
                    "one_call_at_a_time": {"value": total_pix / e2e_call_s / 1e6, "ms_per_step": 1e3 * e2e_call_s / args.steps,
                                           "api": "zw_encode_webp_batch, one blocking call per step (chunks pipelined inside the call)"}},
            "gpu_launches": R["launches"] + (P["launches"] if P else 0),
            "clocks": R["clocks"],
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "dominant_stage": dom,
            "symbols_per_px": R["symbols"] / pix, "bytes_per_px": bytes_per_px,
            "kernel_wall_ms_per_step": 1e3 * wall_kernel_s / args.steps,
            "parity": parity,
        }
        # Rooflines.  The dominant kernel (k_search<2>, the pass-2 luma mode search) is bound by the integer issue rate,
        # not by HBM: `roofline` is its integer figure -- algorithmic int-ops = primitive invocations counted by the
        # instrumented oracle on sample images of this workload x fixed per-primitive costs (tests/oracle_lib.py OP_COST,
        # SURVEY.md 8(d)) against the integer issue peak measured in this run (zw_measure_int_peak).  The HBM view the
        # contract defines is kept beside it as roofline_hbm; roofline_yuv is the one HBM-bound stage.
        hbm = {"bound": "hbm", "kernel": "k_search<2> (pass-2 mode search + transform)", "achieved": pix * SEARCH_BYTES_PER_PX / p2_s / 1e9,
               "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": pix * SEARCH_BYTES_PER_PX / p2_s / 1e9 / peaks["hbm_gbs"], "traffic": traffic,
               "traffic_source": "profiles/ (committed ncu --set full capture, per launch)", "peak_source": how, "ms_per_launch": 1e3 * p2_s,
               "note": "this kernel is integer-issue bound: see `roofline`"}
        try:
            import oracle_lib as O
            ops = {}
            n_s = 4
            for i in range(n_s):
                _, o = O.count_ops(imgs[i * (n // n_s)], QUALITY, METHOD)
                for k, v in o.items():
                    ops[k] = ops.get(k, 0.0) + v / (n_s * W * H)
            all_s = dev_s / args.steps
            line["roofline"] = {
                "bound": "int-issue", "kernel": "k_search<2> (pass-2 luma mode search + final transform + trellis)", "unit": "Tint-op/s",
                "achieved": ops["pass2_luma"] * pix / p2_s / 1e12, "peak": int_peak / 1e12, "frac": ops["pass2_luma"] * pix / p2_s / int_peak,
                "traffic": traffic, "ms_per_launch": 1e3 * p2_s,
                "peak_source": "zw_measure_int_peak in this run (IMAD + LOP3/IADD3 chains at full occupancy); MEASURED_PEAKS.json has no integer peak",
                "ops_per_px": ops,
                "ncu": {k: ps2.get(k) for k in ("alu_pipe_pct", "fma_pipe_pct", "issue_slots_busy_pct", "warp_instructions", "threads_per_instruction") if k in ps2},
                "pass1_luma": {"kernel": "k_search<1>", "achieved": ops["pass1_luma"] * pix / p1_s / 1e12, "frac": ops["pass1_luma"] * pix / p1_s / int_peak,
                               "ms_per_launch": 1e3 * p1_s},
                "whole_step": {"achieved": sum(ops.values()) * pix / all_s / 1e12, "frac": sum(ops.values()) * pix / all_s / int_peak},
                "note": "1 algorithmic op counted as 1 thread-instruction slot; ops counted by the oracle on %d images of the batch" % n_s}
        except Exception as e:  # never fail the bench on the informative figure
            line["roofline"] = dict(hbm, error="int-op count failed: %s" % e)
        line["roofline_hbm"] = hbm
        line["roofline_yuv"] = {"bound": "hbm", "kernel": "k_yuv", "achieved": pix * 4.5 / yuv_s / 1e9, "peak": peaks["hbm_gbs"],
                                "unit": "GB/s", "frac": pix * 4.5 / yuv_s / 1e9 / peaks["hbm_gbs"], "ms_per_launch": 1e3 * yuv_s, "traffic": traffic_yuv}
        lib = None
        if not args.no_cpu and world == 1:
            lib = native_oracle()
            ns = max(cores * 32, 256)  # ~5-10 s of all-core CPU work
            line["cpu_baseline"] = cpu_baseline(host.numpy()[:ns], lib, cores, "of the %d synthetic images (oracle -O3 -march=native)" % n)
        if P:
            p_total = pix * world * args.steps
            photo = {"workload": "%d distinct 768x512 crops of the reference's test photographs (tests/golden/photos), q75 m4" % n,
                     "value": p_total / p_dev_s / 1e6, "ms_per_step": 1e3 * p_dev_s / args.steps,
                     "e2e": {"value": p_total / p_e2e_s / 1e6, "ms_per_step": 1e3 * p_e2e_s / args.steps, "h2d_bytes_per_step": P["h2d"] // args.steps,
                             "d2h_bytes_per_step": P["d2h"] // args.steps, "frac_of_kernel_only": (p_total / p_e2e_s) / (p_total / p_dev_s),
                             "one_call_at_a_time": {"value": p_total / p_call_s / 1e6, "ms_per_step": 1e3 * p_call_s / args.steps}},
                     "stage_ms": {k: v / args.steps for k, v in P["stage_ms"].items()},
                     "symbols_per_px": P["symbols"] / pix, "bytes_per_px": sum(len(o) for o in P["outs"]) / pix}
            be = sum(P["stage_ms"][k] for k in ("token_ms", "boolcode_ms", "assemble_ms")) + 0.0
            photo["back_end_share"] = be / sum(P["stage_ms"].values())
            # EVERY image of the batch against the multi-threaded oracle (after the timed regions)
            import oracle_lib as O
            t0 = time.perf_counter()
            ref = O.encode_batch_mt(photo_host.numpy(), QUALITY, METHOD, threads=cores, L=lib)
            dt = time.perf_counter() - t0
            bad = [i for i in range(n) if not (P["outs"][i] == ref[i] and P["outs_resident"][i] == ref[i] and P["outs_call"][i] == ref[i])]
            photo["parity"] = {"checked": n, "identical": n - len(bad), "against": "CPU port (oracle/), every image, three GPU entry points each"}
            photo["cpu_baseline"] = {"value": n * W * H / dt / 1e6, "unit": "MPix/s", "cores": cores, "kind": "port",
                                     "sample": "all %d photo crops, one image per thread on %d threads (incl. copying the files out)" % (n, cores)}
            line["photo"] = photo
            assert not bad, "photo workload: images %s differ from the oracle" % bad[:8]
        if not args.no_verify:
            try:
                wl = [("synthetic", imgs)] + ([("photo", pimgs)] if P else [])
                line["verify"] = verify_leg(Z, ctx, wl, params)
            except AssertionError:
                raise
            except Exception as e:
                line["verify"] = {"error": str(e)}
        if not args.no_other and world == 1:
            try:
                wl = [("synthetic", imgs)] + ([("photo", pimgs)] if P else [])
                line["lossless"] = lossless_leg(Z, ctx, wl, lib or native_oracle(), cores)
            except AssertionError:
                raise
            except Exception as e:
                line["lossless"] = {"error": str(e)}
        if not args.no_other and world == 1:
            try:
                line["other_configs"] = other_configs(Z, torch, ctx, pipe, lib or native_oracle(), cores, world)
            except AssertionError:
                raise
            except Exception as e:
                line["other_configs"] = {"error": str(e)}
        if not args.no_other:
            try:
                nd = torch.cuda.device_count() if world == 1 else world
                nd = min(nd, args.gpus) if args.gpus > 0 else nd
                line["config5_multi"] = config5_multi(Z, torch, max(1, nd))
            except AssertionError:
                raise
            except Exception as e:
                line["config5_multi"] = {"error": str(e)}
        print(json.dumps(line))
    if dist is not None:
        torch.cuda.synchronize()
        dist.barrier(group=tail_group)
        dist.destroy_process_group()
    pipe.close()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
