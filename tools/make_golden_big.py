#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_big_cases.json: SHA-256 of the oracle's palette and output at the sizes that
are BENCHMARKED (BASELINE.json configs[0..4]), which the oracle cannot redo inside a GPU test run (a 4K CIELAB
image costs it about 1.5 minutes, an 8192x8192 one 10+). The GPU tests (tests/test_gpu_golden_big.py) and
bench.py compare hashes. Like oracle_cases.json these pin the oracle, not the Java reference (no JVM here).

    python tools/make_golden_big.py [--jobs N] [--only SUBSTR] [--list]

Results are merged into the existing file case by case (keyed by `name`), so the run can be interrupted.
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

PATH = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "oracle_big_cases.json"))
IMG_SEED0, RNG_SEED0 = 0x5EED0000, 0xC0FFEE   # bench.py: image i of rank r uses IMG_SEED0 + r * n + i / RNG_SEED0 + r * n + i


def build_image(c):
    """The input of a case. `shape` extras: "q3" = rows [0, q3_rows) flat `q3_color` (more than 2^24 pixels in one
    histogram bin: the float count saturates, PnnQuantizer.java:153 / PnnLABQuantizer.java:154)."""
    from nquant_android_b200.synth import make_image
    img = make_image(c["w"], c["h"], c["cls"], c["alpha"], seed=c["img_seed"])
    if c.get("q3_rows"):
        img = img.copy()
        img[: c["q3_rows"] * c["w"]] = np.uint32(c["q3_color"])
    return img


def cases():
    out = []

    def add(name, kind, cls, alpha, w, h, k, dither, idx=0, **extra):
        out.append(dict(name=name, kind=kind, cls=cls, alpha=alpha, w=w, h=h, k=k, dither=dither,
                        img_seed=IMG_SEED0 + idx, seed=RNG_SEED0 + idx, **extra))

    for cls in ("noisy", "smooth", "rand"):
        add(f"config0_512_{cls}", 0, cls, "opaque", 512, 512, 256, 1)
    add("config1_1080p_lab_noisy", 1, "noisy", "opaque", 1920, 1080, 256, 1)
    add("config2_4k_rgb16_semi", 0, "noisy", "semi", 3840, 2160, 16, 1)
    for idx in (0, 591, 1023):
        add(f"config3_4k_lab_img{idx}", 1, "noisy", "opaque", 3840, 2160, 256, 1, idx=idx)
    for kind in (0, 1):
        q = "lab" if kind else "rgb"
        for k in (2, 16, 64, 256):
            for d in (1, 0):
                add(f"config4_8192_{q}_{k}_{'on' if d else 'off'}", kind, "noisy", "opaque", 8192, 8192, k, d)
        # Q3: 2304 rows x 8192 = 18.9 M pixels of one colour > 2^24
        add(f"q3_8192_{q}_256_on", kind, "noisy", "opaque", 8192, 8192, 256, 1, q3_rows=2304, q3_color=0xFF3C78B4)
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype("<u4").tobytes()).hexdigest()


def run_case(c):
    from oracle import pyoracle as o
    img = build_image(c)
    t0 = time.time()
    try:
        r = o.convert(c["kind"], img, c["w"], c["h"], c["k"], bool(c["dither"]), seed=c["seed"], trace=False)
    except RuntimeError as ex:   # the reference throws here (e.g. ColorUtils.setAlphaComponent, CIELABConvertor.java:79)
        return dict(c, input_sha=sha(img), error=str(ex), oracle_seconds=round(time.time() - t0, 1))
    dt = time.time() - t0
    return dict(c, input_sha=sha(img), palette_sha=sha(r.palette), output_sha=sha(r.out), palette_len=int(len(r.palette)),
                palette=[int(v) for v in r.palette], rng_draws=int(r.scalars["rng_draws"]), maxbins=int(r.scalars["maxbins"]),
                oracle_seconds=round(dt, 1))


def load():
    try:
        return {c["name"]: c for c in json.load(open(PATH))}
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--only", default="")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    from oracle import pyoracle as o
    o.build()
    have = load()
    todo = [c for c in cases() if a.only in c["name"] and (a.force or c["name"] not in have)]
    if a.list:
        for c in cases():
            print(c["name"], "done" if c["name"] in have else "todo")
        return
    todo.sort(key=lambda c: c["w"] * c["h"] * (3 if c["kind"] else 1))   # quick ones first
    with mp.get_context("fork").Pool(a.jobs) as pool:
        for r in pool.imap_unordered(run_case, todo):
            have = load()
            have[r["name"]] = r
            order = [c["name"] for c in cases()]
            json.dump(sorted(have.values(), key=lambda c: order.index(c["name"]) if c["name"] in order else 1 << 30), open(PATH, "w"), indent=1)
            print(f"{r['name']}: {r['oracle_seconds']} s, " + (f"throws: {r['error']}" if "error" in r else f"palette {r['palette_len']}, out {r['output_sha'][:12]}"), flush=True)


if __name__ == "__main__":
    main()
