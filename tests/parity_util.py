"""Stage-by-stage comparison of the CUDA path (through the C ABI) with the CPU oracle."""
import numpy as np

import oracle_lib as O

# stage name -> dtype, in pipeline order (first mismatch localises the bug)
STAGE_ORDER = [("YUV_Y", np.uint8), ("YUV_U", np.uint8), ("YUV_V", np.uint8), ("ALPHA", np.uint8), ("ALPHA_HIST", np.uint32),
               ("SEG_MAP256", np.uint8), ("SEG_CENTERS", np.uint8), ("SEG_MID", np.int32), ("SEG_QIDX", np.uint8),
               ("SEG_MAP", np.uint8), ("SEG_TREE_PROBS", np.uint8), ("SEG_UPDATE_MAP", np.uint8),
               ("P1MB", O.MB_DTYPE), ("STATS", np.uint32), ("PROBS", np.uint8), ("SKIP_PROB", np.uint8), ("LCOST", np.uint16),
               ("P2MB", O.MB_DTYPE), ("HDR_TOKENS", np.uint16), ("TOK_TOKENS", np.uint16), ("PART0", np.uint8), ("PART1", np.uint8), ("VP8", np.uint8)]


def describe_mb_mismatch(name, a, b, mbw):
    msgs = []
    for f in a.dtype.names:
        neq = a[f] != b[f]
        if neq.ndim > 1:
            neq = neq.reshape(neq.shape[0], -1).any(axis=1)
        idx = np.nonzero(neq)[0]
        if idx.size:
            i = int(idx[0])
            msgs.append("%s.%s: %d MBs differ, first at mb %d (x=%d,y=%d): gpu=%s oracle=%s" %
                        (name, f, idx.size, i, i % mbw, i // mbw, np.array2string(a[f][i].reshape(-1)[:40]),
                         np.array2string(b[f][i].reshape(-1)[:40])))
    return "\n".join(msgs)


def compare_stages(ctx, index, dump, mbw, stages=None):
    """Returns list of mismatch descriptions (empty == identical)."""
    out = []
    for name, dt in STAGE_ORDER:
        if stages and name not in stages:
            continue
        if name not in dump:
            continue
        g = ctx.dump_stage(index, name, dt)
        o = dump[name]
        if g.shape != o.shape:
            out.append("%s: shape gpu %s vs oracle %s" % (name, g.shape, o.shape))
            continue
        if dt is O.MB_DTYPE:
            if not (g == o).all():
                out.append(describe_mb_mismatch(name, g, o, mbw))
        else:
            neq = np.nonzero(g != o)[0]
            if neq.size:
                i = int(neq[0])
                out.append("%s: %d of %d entries differ, first at %d: gpu=%s oracle=%s" %
                           (name, neq.size, g.size, i, g[i:i + 8], o[i:i + 8]))
    return out


def check_image(ctx, img, quality, method, params_cls, stop_at_first=True):
    """Encode one image on the GPU and with the oracle; returns (ok, report, gpu_bytes, oracle_bytes)."""
    rc, ref, dump = O.encode(img, quality, method, want_dump=True)
    assert rc == 0
    p = params_cls.lossy(quality)
    p.method = method
    outs, _ = ctx.encode_batch([img], p)
    gpu = outs[0]
    if gpu == ref:
        return True, "", gpu, ref
    mbw = (img.shape[1] + 15) // 16
    rep = compare_stages(ctx, 0, dump, mbw)
    return False, "\n".join(rep[:3] if stop_at_first else rep) or "bytes differ but all stages equal (container?)", gpu, ref
