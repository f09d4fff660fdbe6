"""GPU tests of the streaming (zw_submit / zw_wait / zw_release) and multi-GPU (zw_multi_*) entry points of
the C ABI, and of the device-side stream placement (symbol-arena overflow -> grow and re-run)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import photo_inputs as PI
from image_webp_b200 import synth

pytestmark = pytest.mark.gpu


def _p(q, m):
    import image_webp_b200 as Z
    p = Z.EncoderParams.lossy(q)
    p.method = m
    return p


def test_submit_wait_release_raw_c_abi():
    """Three batches in flight on one context; views point into the slots' pinned arenas; both container forms."""
    import image_webp_b200 as Z
    from image_webp_b200 import _lib
    L = _lib.load()
    ctx = Z.Context(0, depth=3)
    try:
        batches = [[synth.photo_like(160 + 16 * k, 96 + 16 * j, 7 * k + j) for j in range(3)] for k in range(3)]
        keep, tickets = [], []
        for b in batches:
            arr, kp = Z.Context._as_images(b, Z.ColorType.Rgb8)
            keep.append((arr, kp))
            t = C.c_int(-1)
            assert L.zw_submit(ctx.h, arr, len(b), 75, 4, C.byref(t)) == 0
            tickets.append(t.value)
        assert sorted(tickets) == [0, 1, 2]
        arr, kp = Z.Context._as_images(batches[0], Z.ColorType.Rgb8)
        t = C.c_int(-1)
        assert L.zw_submit(ctx.h, arr, 3, 75, 4, C.byref(t)) == 7  # ZW_ERR_BUSY: every slot taken
        for b, tk in zip(batches, tickets):
            view = _lib.ZwBatchView()
            tm = _lib.ZwTiming()
            assert L.zw_wait(ctx.h, tk, 1, C.byref(view), C.byref(tm)) == 0
            assert view.n == len(b) and tm.kernel_launches >= 17
            files = [C.string_at(view.arena + view.offsets[i], view.lens[i]) for i in range(view.n)]
            assert L.zw_wait(ctx.h, tk, 0, C.byref(view), None) == 0  # same ticket, raw VP8 payloads
            raws = [C.string_at(view.arena + view.offsets[i], view.lens[i]) for i in range(view.n)]
            for img, f, r in zip(b, files, raws):
                rc, ref, _ = O.encode(img, 75, 4)
                assert f == ref and r == ref[20:20 + len(r)] and len(r) == int.from_bytes(ref[16:20], "little")
            assert L.zw_release(ctx.h, tk) == 0
        assert L.zw_wait(ctx.h, tickets[0], 1, None, None) == 6  # released: ZW_ERR_NOT_STAGED
        assert L.zw_submit(ctx.h, arr, 3, 75, 4, C.byref(t)) == 0 and L.zw_release(ctx.h, t.value) == 0  # abandon
    finally:
        ctx.close()


def test_pipeline_wrapper_many_batches_in_order():
    import image_webp_b200 as Z
    batches = [list(PI.batch(4, 256, 192, first=4 * k)) for k in range(7)]
    with Z.BatchPipeline(0, depth=2) as pipe:
        got = pipe.encode_batches(batches, _p(75, 4))
    for b, outs in zip(batches, got):
        ref = O.encode_batch_mt(np.stack(b), 75, 4)
        assert outs == ref


def test_symbol_arena_overflow_grows_and_reruns():
    """q100 noise carries ~10 symbols per pixel, far above the 3 per pixel the token arena is first sized for: the
    device flags the overflow, the host grows the arena and re-runs the emit / code / assemble kernels."""
    import image_webp_b200 as Z
    ctx = Z.Context(0)
    try:
        imgs = [synth.noise(256, 256, 11 + i) for i in range(3)]
        outs, t = ctx.encode_batch(imgs, _p(100, 4))
        assert t["symbols"] > 6 * t["pixels"]
        for im, o in zip(imgs, outs):
            assert o == O.encode(im, 100, 4)[1]
        outs2, _ = ctx.encode_batch(imgs, _p(100, 4))  # second time: the learned estimate fits
        assert outs2 == outs
    finally:
        ctx.close()


def test_large_batch_is_chunked_and_pipelined_identically():
    import image_webp_b200 as Z
    b = PI.batch(40, 512, 384)  # ~7.9 Mpx: one chunk by default
    ref = O.encode_batch_mt(b, 75, 4)
    small = Z.Context(0, max_device_bytes=40 << 20)  # forces ~8 chunks over 3 slots
    try:
        outs, t = small.encode_batch(list(b), _p(75, 4))
        assert outs == ref and t["kernel_launches"] >= 5 * 17
    finally:
        small.close()


def test_contiguous_and_scattered_inputs_give_the_same_bytes():
    import image_webp_b200 as Z
    ctx = Z.Context(0)
    try:
        b = PI.batch(6, 256, 256)  # one contiguous array: coalesced H2D
        a, _ = ctx.encode_batch(list(b), _p(75, 4))
        scattered = [np.array(x) for x in b]
        s, _ = ctx.encode_batch(scattered, _p(75, 4))
        mixed = [b[0], scattered[1], b[2], b[3], scattered[4], b[5]]
        m, _ = ctx.encode_batch(mixed, _p(75, 4))
        assert a == s == m == O.encode_batch_mt(b, 75, 4)
    finally:
        ctx.close()


def test_multi_gpu_entry_on_all_visible_devices():
    """zw_multi_encode shards by image over every visible GPU (1 on the single-GPU box, more under --gpus N);
    a device list that repeats device 0 exercises the slicing + gather on one GPU as well."""
    import torch
    import image_webp_b200 as Z
    from image_webp_b200 import shard
    n_dev = torch.cuda.device_count()
    imgs = [synth.photo_like(128 + 16 * (i % 4), 96 + 16 * (i % 3), 300 + i) for i in range(11)]
    ref = [O.encode(im, 75, 4)[1] for im in imgs]
    for devices in ([0, 0, 0], list(range(n_dev))):
        mc = Z.MultiContext(devices)
        try:
            outs, tm = mc.encode_batch(imgs, _p(75, 4))
            assert outs == ref and len(tm) == len(devices)
            assert sum(t["pixels"] for t in tm) == sum(im.shape[0] * im.shape[1] for im in imgs)
        finally:
            mc.close()
    assert shard.encode_batch_sharded(imgs, _p(75, 4), list(range(n_dev))) == ref
