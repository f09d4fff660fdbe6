"""Image sharding across GPUs: images are independent, so a batch is split into
contiguous slices of ceil(n/G) images, one slice per GPU, no collective on the data path; the
finished bitstreams are gathered on the host in image order."""


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


_multi = {}


def encode_batch_sharded(images, params, devices, color=None, container=True):
    """The library's multi-GPU entry (zw_multi_encode): one host thread + context per GPU inside the C ABI,
    contiguous slices of ceil(n / G) images, results in image order."""
    from .encoder import ColorType, MultiContext
    color = color or ColorType.Rgb8
    key = tuple(devices)
    if key not in _multi:
        _multi[key] = MultiContext(list(devices))
    return _multi[key].encode_batch(images, params, color, container)[0]
