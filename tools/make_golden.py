#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_cases.json: SHA-256 of the oracle's palette and output for a set of
small synthetic cases (shared-math mode). These pin the ORACLE against regressions and give the GPU
tests a fixture that travels; they are not outputs of the Java reference (no JVM exists here)."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from oracle import pyoracle as o
from nquant_android_b200.synth import make_image

CASES = []
for kind in (0, 1):
    for cls, alpha, w, h, k, d in [
        ("noisy", "opaque", 96, 64, 256, 1), ("smooth", "opaque", 96, 64, 256, 1), ("rand", "opaque", 64, 64, 64, 1),
        ("noisy", "semi", 96, 64, 16, 1), ("noisy", "transparent", 96, 64, 64, 1), ("noisy", "opaque", 96, 64, 256, 0),
        ("smooth", "opaque", 50, 37, 2, 1), ("smooth", "transparent", 50, 37, 2, 1), ("noisy", "opaque", 33, 70, 4, 1),
        ("noisy", "opaque", 96, 64, 16, 0), ("smooth", "semi", 64, 48, 256, 1), ("noisy", "opaque", 1, 97, 8, 1),
        ("noisy", "opaque", 97, 1, 8, 1),
    ]:
        CASES.append(dict(kind=kind, cls=cls, alpha=alpha, w=w, h=h, k=k, dither=d, seed=0xC0FFEE))

def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype("<u4").tobytes()).hexdigest()

out = []
for c in CASES:
    img = make_image(c["w"], c["h"], c["cls"], c["alpha"])
    if c["kind"] == 1 and not c["dither"] and c["k"] > 32:
        continue   # BlueNoise second pass of the LAB quantizer: not covered by the CUDA path yet
    r = o.convert(c["kind"], img, c["w"], c["h"], c["k"], bool(c["dither"]), seed=c["seed"], trace=False)
    out.append(dict(c, input_sha=sha(img), palette_sha=sha(r.palette), output_sha=sha(r.out), palette_len=int(len(r.palette))))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "oracle_cases.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", len(out), "cases to", os.path.normpath(path))
