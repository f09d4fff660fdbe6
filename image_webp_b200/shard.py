"""Image sharding across GPUs: images are independent, so a batch is split into
contiguous slices of ceil(n/G) images, one slice per GPU, no collective on the data path; the
finished bitstreams are gathered on the host in image order."""
import threading


def shard_range(n, rank, world):
    """[begin, end) of the images rank `rank` of `world` owns."""
    if world <= 0 or rank < 0 or rank >= world:
        raise ValueError("bad rank/world")
    per = (n + world - 1) // world
    b = min(n, rank * per)
    return b, min(n, b + per)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_in_order(shards):
    """Concatenate per-rank result lists back into image order."""
    out = []
    for s in shards:
        out.extend(s)
    return out


def encode_batch_sharded(images, params, devices, color=None, container=True):
    """One host thread + one context per GPU (contexts are independent)."""
    from .encoder import ColorType, default_context
    color = color or ColorType.Rgb8
    world = len(devices)
    results = [None] * world
    errors = [None] * world

    def work(r):
        try:
            b, e = shard_range(len(images), r, world)
            results[r] = default_context(devices[r]).encode_batch(images[b:e], params, color, container)[0] if e > b else []
        except Exception as ex:  # surfaced to the caller below
            errors[r] = ex

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for ex in errors:
        if ex is not None:
            raise ex
    return gather_in_order(results)
