"""Smallest possible GPU check of the speculative dither (no torch, no pytest): one 256x256 PnnLABQuantizer image,
speculative path on, compared with the oracle; then the same through the serial path with timings.
    python tools/spec_gpu_probe.py [W H SEG WARM [BATCH]]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nquant_android_b200.quantizer import Context  # noqa: E402
from nquant_android_b200.synth import make_image  # noqa: E402
from oracle import pyoracle  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    w, h, seg, warm = (a + [256, 256, 4096, 1024][len(a):])[:4]
    batch = a[4] if len(a) > 4 else 1
    imgs = np.stack([make_image(w, h, "noisy", "opaque", seed=0x5EED0000 + i) for i in range(batch)])
    seeds = [0xC0FFEE + i for i in range(batch)]
    ref = pyoracle.convert(1, imgs[0], w, h, 256, True, seed=seeds[0], trace=False) if (w * h <= 1 << 21 and not os.environ.get("NQ_PROBE_NOORACLE")) else None
    ctx = Context(0)
    for spec in (1, 0, 1):
        ctx.set_spec_dither(bool(spec), seg, warm)
        t0 = time.time()
        out, pal, plen, _ = ctx.convert_batch(1, imgs, w, h, 256, True, seeds=seeds)
        dt = time.time() - t0
        ok = None if ref is None else bool(np.array_equal(out[0], ref.out) and np.array_equal(pal[0, :plen[0]], ref.palette))
        print(f"spec={spec} {batch} x {w}x{h}: {dt * 1e3:.1f} ms wall, matches oracle: {ok}, stats {ctx.spec_stats()}, "
              f"draws {ctx.image_info(0)['rng_draws']}" + ("" if ref is None else f" (oracle {ref.scalars['rng_draws']})"), flush=True)
        if spec == 1:
            keep = out.copy()
        else:
            print("spec output == serial output:", bool(np.array_equal(out, keep)),
                  "differing pixels per image:", [int((out[i] != keep[i]).sum()) for i in range(batch)], flush=True)
    ms = ctx.stage_times(reset=True) if hasattr(ctx, "stage_times") else None
    print("stage ms:", ms)


if __name__ == "__main__":
    main()
