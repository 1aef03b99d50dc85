"""Synthetic ARGB test/bench images. Pure integer arithmetic so the numpy
generator here and the CUDA generator (csrc/nq_synth.cuh) produce identical bytes.

  h(seed, idx, ch) = mix64(seed ^ ((idx*4 + ch) * 0x9E3779B97F4A7C15))
  base gradient   R = 255*x//(W-1), G = 255*y//(H-1), B = 255*(x+y)//(W+H-2)
  classes         smooth: amp 2, noisy: amp 32 (noise uniform in [-amp, amp], clamped), rand: uniform bytes,
                  few (numpy only): rand posterised to 2 levels of r, 3 of g, 2 of b = at most 12 colours
  alpha modes     opaque: 255; transparent: 255 with A=0 in the top-left (W/8 x H/8) block;
                  semi: A = 255*(W-1-x)//(W-1) with A=0 in the same block
"""
import numpy as np

CLASSES = {"smooth": 0, "noisy": 1, "rand": 2, "few": 3}
ALPHA = {"opaque": 0, "transparent": 1, "semi": 2}
AMP = {0: 2, 1: 32, 2: 0}
GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def make_image(width, height, cls="noisy", alpha="opaque", seed=0x5EED0000):
    """Returns a (height*width,) uint32 array of non-premultiplied 0xAARRGGBB pixels."""
    kls = CLASSES[cls] if isinstance(cls, str) else int(cls)
    amode = ALPHA[alpha] if isinstance(alpha, str) else int(alpha)
    n = width * height
    idx = np.arange(n, dtype=np.uint64)
    x = (idx % np.uint64(width)).astype(np.int64)
    y = (idx // np.uint64(width)).astype(np.int64)
    with np.errstate(over="ignore"):
        hs = [_mix64(np.uint64(seed) ^ ((idx * np.uint64(4) + np.uint64(ch)) * GOLDEN)) for ch in range(3)]
    if kls == 3:   # icon-like: a handful of flat colours (exercises PnnLABQuantizer's few-colours shortcut)
        r = ((hs[0] & np.uint64(1)).astype(np.int64)) * 200 + 30
        g = ((hs[1] % np.uint64(3)).astype(np.int64)) * 100 + 20
        b = ((hs[2] & np.uint64(1)).astype(np.int64)) * 180 + 40
    elif kls == 2:
        r, g, b = [(h & np.uint64(0xFF)).astype(np.int64) for h in hs]
    else:
        amp = AMP[kls]
        span = np.uint64(2 * amp + 1)
        base = [255 * x // max(width - 1, 1), 255 * y // max(height - 1, 1), 255 * (x + y) // max(width + height - 2, 1)]
        r, g, b = [np.clip(bv + (h % span).astype(np.int64) - amp, 0, 255) for bv, h in zip(base, hs)]
    if amode == 0:
        a = np.full(n, 255, dtype=np.int64)
    else:
        a = np.full(n, 255, dtype=np.int64) if amode == 1 else 255 * (width - 1 - x) // max(width - 1, 1)
        a = np.where((x < width // 8) & (y < height // 8), 0, a)
    return ((a << 24) | (r << 16) | (g << 8) | b).astype(np.uint32)
