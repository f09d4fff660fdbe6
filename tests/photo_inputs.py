"""Real-photograph inputs for the parity tests and the bench's photo workload.

tests/golden/photos/{3,4,5}.png are the reference's own test photographs
(/root/reference/tests/reference/gallery1/{3,4,5}.png: 1280x720, 1024x772, 1024x752 RGB, the PNG
decodes of Google's WebP gallery, tests/CREDITS.md) -- test vectors, not source.  SURVEY.md 8(d)
config 1(i) names the (256,104) 768x512 crop of 3.png; the reference's own bench encodes a
768x512 Kodak photograph (benches/profile_encode.rs:33)."""
import functools
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PHOTO_DIR = os.path.join(HERE, "golden", "photos")
NAMES = ("3", "4", "5")


@functools.lru_cache(maxsize=None)
def photo(name):
    """Full RGB photograph as a uint8 array [h,w,3]."""
    from PIL import Image
    im = Image.open(os.path.join(PHOTO_DIR, "%s.png" % name)).convert("RGB")
    return np.ascontiguousarray(np.asarray(im, dtype=np.uint8))


def crop(name, x, y, w=768, h=512):
    p = photo(str(name))
    assert 0 <= x and 0 <= y and x + w <= p.shape[1] and y + h <= p.shape[0]
    return np.ascontiguousarray(p[y:y + h, x:x + w])


def survey_crop():
    """SURVEY.md 8(d) config 1(i)."""
    return crop("3", 256, 104)


def crop_origin(i, w=768, h=512):
    """Photo and origin of the i-th crop of the sliding-offset batch: the three photographs in turn, offsets
    stepping by (37, 23) modulo the free range (distinct for every i < 1024 at 768x512)."""
    name = NAMES[i % 3]
    j = i // 3
    p = photo(name)
    fx, fy = p.shape[1] - w + 1, p.shape[0] - h + 1
    return name, (j * 37) % fx, (j * 23) % fy


def batch(n, w=768, h=512, out=None, first=0):
    """n distinct crops stacked [n,h,w,3] (written into `out` when given, e.g. a pinned buffer)."""
    if out is None:
        out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        name, x, y = crop_origin(first + i, w, h)
        out[i] = photo(name)[y:y + h, x:x + w]
    return out
