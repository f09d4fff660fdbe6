#!/usr/bin/env python3
"""bench.py -- lossy encode MPix/s (q75 m4, byte-identical) on N B200s, beside the CPU baseline.

Workload (BASELINE.json configs[1]): a batch of 1024 x 768x512 RGB images, quality 75, method 4,
synthetic "photo-like" content (image_webp_b200/synth.py).  One step = one encode of the whole
batch.  Images are independent, so N GPUs each encode their own batch of 1024 (weak scaling, no
collective on the data path); the metric is Σ pixels of all ranks / max-over-ranks time.

  value : kernel-only, inputs already resident in HBM (zw_encode_resident), CUDA-event time
  e2e   : zw_encode_webp_batch from pinned HOST buffers: H2D + kernels + D2H + host RIFF assembly
  --impl reference : the CPU oracle (a C++ restatement of the reference; the Rust reference cannot
                     be built in this image) on all host cores, one image per thread.
"""
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
# algorithmic HBM bytes per pixel of the dominant kernel (pass-2 luma search k_search<2>): read the Y
# plane (1 B/px) + write the 32-byte header and 17 luma blocks (544 B) of one macroblock record per
# 256 px (2.25 B/px).  DESIGN.md §Measurement.
SEARCH_BYTES_PER_PX = 1.0 + 576.0 / 256.0  # luma kernel: Y plane read + header and luma part of the record written


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
    return L


def cpu_run(imgs, threads, lib):
    """Encode imgs ([n,h,w,3] uint8) with the oracle, one image per thread at a time.  Returns seconds."""
    n, h, w, _ = imgs.shape
    t0 = time.perf_counter()
    lib.zwo_encode_batch_mt(imgs.ctypes.data, n, w, h, QUALITY, METHOD, threads)
    return time.perf_counter() - t0


def run_reference(args):
    rank, _, world = rank_env()
    if rank != 0:
        return 0
    import numpy as np
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
    line = {"metric": "lossy encode MPix/s (q75 m4, byte-identical)", "value": value, "unit": "MPix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tsum / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD % 1024, "quality": QUALITY, "method": METHOD,
                       "note": "bounded sample of the workload per step; CPU only"},
            "cpu_baseline": {"value": value, "unit": "MPix/s", "cores": cores, "kind": "port",
                             "sample": "%d of the 1024 768x512 images per step, one image per thread on %d threads" % (n_sample, cores)},
            "e2e": {"value": value, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--depth", type=int, default=3, help="batches in flight in the end-to-end leg (BatchPipeline depth)")
    ap.add_argument("--check", type=int, default=8, help="images per step byte-compared with the oracle after timing")
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
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    import image_webp_b200 as Z
    from image_webp_b200 import synth
    n = args.batch
    # distinct content per rank; pinned host staging (the e2e leg copies from here every step)
    host = torch.empty((n, H, W, 3), dtype=torch.uint8, pin_memory=True)
    base = synth.batch_photo_like(n, W, H, seed0=1000 * rank)
    host.numpy()[...] = base
    del base
    imgs = [host.numpy()[i] for i in range(n)]
    params = Z.EncoderParams.lossy(QUALITY)
    params.method = METHOD
    ctx = Z.Context(dev)
    pix = n * W * H

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only: inputs resident in HBM -------------------------------------------------
    ctx.stage(imgs)
    for _ in range(args.warmup):
        ctx.encode_resident(params)
    sampler = ClockSampler(dev)
    barrier()
    sampler.start()
    stage_ms = {}
    dev_ms = 0.0
    launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t = ctx.encode_resident(params)
        dev_ms += t["device_total_ms"]
        launches += t["kernel_launches"]
        for k in ("yuv_ms", "analysis_ms", "pass1_ms", "chroma1_ms", "stats_ms", "chroma2_ms", "pass2_ms", "token_ms", "boolcode_ms", "assemble_ms"):
            stage_ms[k] = stage_ms.get(k, 0.0) + t[k]
    barrier()
    wall_kernel_s = time.perf_counter() - t0
    clocks = sampler.stop()
    outs_resident, _ = ctx.download()

    # ---- end to end: host buffers in, .webp bytes out ------------------------------------------
    # (a) one call at a time through Context.encode_batch (zw_encode_webp_batch)
    for _ in range(2):
        ctx.encode_batch(imgs, params)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outs, t = ctx.encode_batch(imgs, params)
    barrier()
    e2e_serial_s = time.perf_counter() - t0
    # (b) the streaming entry point (BatchPipeline, args.depth contexts / host threads): every step still
    #     copies its inputs H2D and its .webp bytes D2H, but under the kernels of the neighbouring step
    pipe = Z.BatchPipeline(dev, depth=args.depth)
    for f in [pipe.submit(imgs, params) for _ in range(args.depth)]:
        f.result()
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    futs = [pipe.submit(imgs, params) for _ in range(args.steps)]
    for f in futs:
        outs, t = f.result()
        h2d += t["h2d_bytes"]; d2h += t["d2h_bytes"]
    barrier()
    e2e_s = time.perf_counter() - t0
    pipe.close()

    int_peak = ctx.measure_int_peak()  # thread-level integer instructions / s (microbenchmark, outside the timed regions)

    # ---- parity spot check (after the timed regions) -----------------------------------------
    parity = None
    if rank == 0 and args.check > 0:
        import oracle_lib as O
        ok = 0
        idx = list(range(0, n, max(1, n // args.check)))[:args.check]
        for i in idx:
            rc, ref, _ = O.encode(imgs[i], QUALITY, METHOD)
            ok += int(rc == 0 and outs[i] == ref and outs_resident[i] == ref)
        parity = {"checked": len(idx), "identical": ok}
        assert ok == len(idx), "GPU output differs from the oracle"

    # ---- reduce over ranks: max time ------------------------------------------------------------
    times = torch.tensor([dev_ms / 1e3, e2e_s, wall_kernel_s, e2e_serial_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, e2e_s, wall_kernel_s, e2e_serial_s = [float(x) for x in times.tolist()]
    total_pix = pix * world * args.steps
    value = total_pix / dev_s / 1e6
    e2e_value = total_pix / e2e_s / 1e6

    if rank == 0:
        peaks, how = measured_peaks()
        dom = max(("pass1_ms", "pass2_ms", "boolcode_ms", "token_ms", "stats_ms", "yuv_ms", "analysis_ms"), key=lambda k: stage_ms[k])
        dom_s = stage_ms["pass2_ms"] / args.steps / 1e3   # k_search<2> alone (CUDA events around that launch)
        achieved = pix * SEARCH_BYTES_PER_PX / dom_s / 1e9
        yuv_s = stage_ms["yuv_ms"] / args.steps / 1e3
        traffic = traffic_yuv = None
        try:  # DRAM bytes per pixel from the committed ncu --set full capture (profiles/r1_traffic.json)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["kernels"]
            for kname, v in tj.items():
                if "k_search<2>" in kname or "k_search<(int)2>" in kname:
                    traffic = v["dram_bytes_per_pixel"] * pix
                if "k_yuv" in kname:
                    traffic_yuv = v["dram_bytes_per_pixel"] * pix
        except Exception:
            pass
        line = {
            "metric": "lossy encode MPix/s (q75 m4, byte-identical)", "value": value, "unit": "MPix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": WORKLOAD % n, "images_per_gpu": n, "quality": QUALITY, "method": METHOD,
                       "l2": "inputs (%.2f GB per step) larger than L2" % (n * W * H * 3 / 1e9), "timing": "cuda events on the library stream, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "MPix/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "BatchPipeline(depth=%d).submit -> zw_encode_webp_batch, pinned host RGB in, .webp bytes out" % args.depth,
                    "one_call_at_a_time": {"value": total_pix / e2e_serial_s / 1e6, "ms_per_step": 1e3 * e2e_serial_s / args.steps,
                                           "api": "Context.encode_batch -> zw_encode_webp_batch"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_search<2> (pass-2 mode search + transform)", "achieved": achieved,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                         "peak_source": how, "ms_per_launch": 1e3 * dom_s,
                         "note": "integer-issue bound, not HBM bound: see DESIGN.md and profiles/ for the ALU pipe figures"},
            "roofline_yuv": {"bound": "hbm", "kernel": "k_yuv", "achieved": pix * 4.5 / yuv_s / 1e9, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": pix * 4.5 / yuv_s / 1e9 / peaks["hbm_gbs"], "ms_per_launch": 1e3 * yuv_s, "traffic": traffic_yuv},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "dominant_stage": dom,
            "kernel_wall_ms_per_step": 1e3 * wall_kernel_s / args.steps,
            "parity": parity,
        }
        # Integer roofline of the mode search (north_star: "ALU/IMAD pipe utilisation against the sm_100a integer issue
        # peak"): algorithmic int-ops = primitive invocations counted by the instrumented oracle on sample images of this
        # workload x fixed per-primitive costs (tests/oracle_lib.py OP_COST), per pass and plane.
        try:
            import oracle_lib as O
            ops = {}
            n_s = 4
            for i in range(n_s):
                _, o = O.count_ops(imgs[i * (n // n_s)], QUALITY, METHOD)
                for k, v in o.items():
                    ops[k] = ops.get(k, 0.0) + v / (n_s * W * H)
            p2_s = stage_ms["pass2_ms"] / args.steps / 1e3
            p1_s = stage_ms["pass1_ms"] / args.steps / 1e3
            all_s = dev_s / args.steps
            line["roofline_int"] = {
                "bound": "int-issue", "unit": "Tint-op/s", "peak": int_peak / 1e12, "peak_source": "zw_measure_int_peak (IMAD + LOP3/IADD3 chains, this run)",
                "ops_per_px": ops,
                "kernel": "k_search<2> (pass-2 luma)", "achieved": ops["pass2_luma"] * pix / p2_s / 1e12,
                "frac": ops["pass2_luma"] * pix / p2_s / int_peak,
                "pass1_luma": {"achieved": ops["pass1_luma"] * pix / p1_s / 1e12, "frac": ops["pass1_luma"] * pix / p1_s / int_peak},
                "whole_step": {"achieved": sum(ops.values()) * pix / all_s / 1e12, "frac": sum(ops.values()) * pix / all_s / int_peak},
                "note": "1 algorithmic op counted as 1 instruction slot; sample = %d images of the batch" % n_s}
        except Exception as e:  # the figure is informative; never fail the bench on it
            line["roofline_int"] = {"error": str(e)}
        if not args.no_cpu and world == 1:
            lib = native_oracle()
            cores = os.cpu_count() or 1
            ns = max(cores * 32, 256)  # ~5-10 s of all-core CPU work
            sample = host.numpy()[:ns]
            cpu_run(sample[:cores], cores, lib)
            ts = cpu_run(sample, cores, lib)
            t1 = cpu_run(sample[:4], 1, lib)
            line["cpu_baseline"] = {"value": ns * W * H / ts / 1e6, "unit": "MPix/s", "cores": cores, "kind": "port",
                                    "single_thread_mpix_s": 4 * W * H / t1 / 1e6,
                                    "sample": "%d of the %d images, one image per thread on %d threads (oracle -O3 -march=native)" % (ns, n, cores)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
