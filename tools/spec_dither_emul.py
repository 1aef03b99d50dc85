"""Runs tools/spec_dither_emul.cpp (CPU emulation of the speculative segment-parallel Gilbert dither,
SURVEY.md section 8f rank 1) over the synthetic input classes and prints a markdown table: rounds until
every segment is validated, redundant work, why validations failed, and whether the assembled output is
bit-identical to the sequential oracle. CPU only; study infrastructure.

    python tools/spec_dither_emul.py [--size 960x540] [--seg 8192] [--warm 2048]
"""
import argparse
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from nquant_android_b200.synth import make_image  # noqa: E402

SO = os.path.join(ROOT, "build", "libnq_spec_emul.so")

CASES = [
    ("PnnLABQuantizer", 1, 256, True, "noisy", "opaque"),
    ("PnnQuantizer", 0, 256, True, "noisy", "opaque"),
    ("PnnLABQuantizer", 1, 256, True, "rand", "opaque"),
    ("PnnQuantizer", 0, 256, True, "rand", "opaque"),
    ("PnnQuantizer", 0, 16, True, "noisy", "semi"),
    ("PnnLABQuantizer", 1, 16, True, "noisy", "opaque"),
    ("PnnQuantizer", 0, 64, True, "noisy", "opaque"),
    ("PnnLABQuantizer", 1, 64, True, "noisy", "opaque"),
    ("PnnQuantizer", 0, 16, True, "noisy", "opaque"),
    ("PnnQuantizer", 0, 2, True, "noisy", "opaque"),
    ("PnnLABQuantizer", 1, 256, True, "noisy", "transparent"),
    ("PnnQuantizer", 0, 256, False, "noisy", "opaque"),
    ("PnnQuantizer", 0, 64, False, "noisy", "opaque"),
]


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(HERE, "spec_dither_emul.cpp")
    dep = os.path.join(ROOT, "oracle", "nq_oracle.cpp")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        fma = ["-mfma"] if " fma " in open("/proc/cpuinfo").read() else []
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w"] + fma +
                              ["-o", SO, src])
    L = ctypes.CDLL(SO)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.nqs_spec_emulate.argtypes = [ci, vp, ci, ci, ci, ci, ctypes.c_uint64, ci, ci, vp]
    return L


def emulate(L, kind, img, w, h, nmax, dither, seg, warm):
    out = np.zeros(16, np.int64)
    img = np.ascontiguousarray(img, dtype=np.uint32)
    rc = L.nqs_spec_emulate(kind, img.ctypes.data, w, h, nmax, int(dither), 0xC0FFEE, seg, warm, out.ctypes.data)
    if rc:
        raise RuntimeError("emulation failed")
    keys = ["rounds", "nseg", "DM", "sorted", "exact", "pixelsRun", "segRuns", "claims", "draws", "failQ", "failD", "failM", "maxRerun", "specReads", "errDependent"]
    return dict(zip(keys, [int(v) for v in out]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="960x540")
    ap.add_argument("--seg", type=int, default=8192)
    ap.add_argument("--warm", type=int, default=2048)
    ap.add_argument("--only", type=int, default=-1)
    a = ap.parse_args()
    w, h = [int(v) for v in a.size.split("x")]
    L = build()
    print(f"segment {a.seg} pixels, warm-up {a.warm} pixels, image {w}x{h}\n")
    print("| quantizer | colours | class/alpha | DITHER_MAX | segments | rounds | segment runs | pixels run / pixels | "
          "failed on Q / D / memo | memo claims | predicted-memo reads | draws | error-dependent lookups | bit-identical |")
    print("|---|---:|---|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|---|")
    for i, (name, kind, nmax, dither, cls, alpha) in enumerate(CASES):
        if a.only >= 0 and i != a.only:
            continue
        img = make_image(w, h, cls, alpha)
        r = emulate(L, kind, img, w, h, nmax, dither, a.seg, a.warm)
        if r["sorted"]:
            print(f"| {name} | {nmax} | {cls}/{alpha} | {r['DM']} | - | - | - | - | PriorityQueue mode: not covered | - | - | - | - | - |", flush=True)
            continue
        print(f"| {name} | {nmax} | {cls}/{alpha} | {r['DM']} | {r['nseg']} | {r['rounds']} | {r['segRuns']} | "
              f"{r['pixelsRun'] / (w * h):.2f} | {r['failQ']} / {r['failD']} / {r['failM']} | {r['claims']} | {r['specReads']} | {r['draws']} | "
              f"{100.0 * r['errDependent'] / (w * h):.1f} % | "
              f"{'yes' if r['exact'] else 'NO'} |", flush=True)


if __name__ == "__main__":
    main()
