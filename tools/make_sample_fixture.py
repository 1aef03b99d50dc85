#!/usr/bin/env python3
"""Makes tests/golden/sample_495x438.png and tests/golden/sample_cases.json from the ONE real image the reference
ships (its demo picture, /root/reference/app/src/main/res/drawable/sample.jpg, converted by the demo as
`new PnnQuantizer(path).convert(256, true)`, MainActivity.java:190-194). /root/reference does not exist on the GPU box,
so the decoded pixels travel as a lossless PNG; the oracle's palette and output for the demo's call and for
PnnLABQuantizer are frozen next to it. The JPEG was decoded here with Pillow (Android's decoder may differ in the last
bit of a few pixels: this is a real-image input, not a claim about the reference's output on a device)."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from PIL import Image
from oracle import pyoracle as o
from nquant_android_b200.imageio import load_argb

SRC = "/root/reference/app/src/main/res/drawable/sample.jpg"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype("<u4").tobytes()).hexdigest()


argb, w, h = load_argb(SRC)
Image.open(SRC).convert("RGB").save(os.path.join(GOLD, f"sample_{w}x{h}.png"), optimize=True)
back, w2, h2 = load_argb(os.path.join(GOLD, f"sample_{w}x{h}.png"))
assert (w2, h2) == (w, h) and np.array_equal(back, argb)
cases = []
for kind, k, d in [(0, 256, 1), (1, 256, 1), (0, 16, 1), (1, 64, 0), (0, 256, 0)]:
    r = o.convert(kind, argb, w, h, k, bool(d), seed=0xC0FFEE, trace=False)
    cases.append(dict(kind=kind, k=k, dither=d, seed=0xC0FFEE, w=w, h=h, input_sha=sha(argb), palette=[int(v) for v in r.palette],
                      output_sha=sha(r.out), rng_draws=int(r.scalars["rng_draws"]), maxbins=int(r.scalars["maxbins"])))
    print(kind, k, d, len(r.palette), cases[-1]["output_sha"][:12])
json.dump(cases, open(os.path.join(GOLD, "sample_cases.json"), "w"), indent=1)
