"""Runs tools/resync_study.cpp over the synthetic input classes: how many pixels after a restart with an
empty error queue does the Gilbert dither's queue become bit-identical to the sequential run's again?
 CPU only; uses the oracle, so it is study/test infrastructure.

    python tools/resync_study.py [--size 960x540] [--starts 24] [--maxlen 300000]
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

SO = os.path.join(ROOT, "build", "libnq_resync_study.so")


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(HERE, "resync_study.cpp")
    dep = os.path.join(ROOT, "oracle", "nq_oracle.cpp")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        fma = ["-mfma"] if " fma " in open("/proc/cpuinfo").read() else []
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w"] + fma +
                              ["-o", SO, src])
    L = ctypes.CDLL(SO)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.nqs_resync_study.argtypes = [ci, vp, ci, ci, ci, ci, ctypes.c_uint64, ci, ci, vp, vp, vp]
    return L


def study(L, kind, img, w, h, nmax, dither, nstarts, maxlen):
    starts = np.zeros(nstarts, np.int32)
    lengths = np.zeros(nstarts, np.int32)
    info = np.zeros(4, np.int32)
    img = np.ascontiguousarray(img, dtype=np.uint32)
    n = L.nqs_resync_study(kind, img.ctypes.data, w, h, nmax, int(dither), 0xC0FFEE, nstarts, maxlen,
                           starts.ctypes.data, lengths.ctypes.data, info.ctypes.data)
    if n < 0:
        raise RuntimeError("study failed")
    return starts[:n], lengths[:n], info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="960x540")
    ap.add_argument("--starts", type=int, default=24)
    ap.add_argument("--maxlen", type=int, default=300000)
    ap.add_argument("--md", action="store_true")
    a = ap.parse_args()
    w, h = [int(v) for v in a.size.split("x")]
    L = build()
    cases = [
        ("PnnLABQuantizer", 1, 256, True, "noisy", "opaque"),
        ("PnnQuantizer", 0, 256, True, "noisy", "opaque"),
        ("PnnLABQuantizer", 1, 256, True, "smooth", "opaque"),
        ("PnnQuantizer", 0, 256, True, "smooth", "opaque"),
        ("PnnLABQuantizer", 1, 256, True, "rand", "opaque"),
        ("PnnQuantizer", 0, 256, True, "rand", "opaque"),
        ("PnnQuantizer", 0, 16, True, "noisy", "semi"),
        ("PnnLABQuantizer", 1, 16, True, "noisy", "opaque"),
        ("PnnQuantizer", 0, 64, True, "noisy", "opaque"),
        ("PnnLABQuantizer", 1, 64, True, "noisy", "opaque"),
        ("PnnQuantizer", 0, 2, True, "noisy", "opaque"),
        ("PnnQuantizer", 0, 256, False, "noisy", "opaque"),
        ("PnnLABQuantizer", 1, 256, False, "noisy", "opaque"),
    ]
    print(f"| quantizer | colours | dither | class/alpha | size | queue (DITHER_MAX, ditherMax) | restarts | re-synchronised | "
          f"median | p90 | max | not within {a.maxlen} |")
    print("|---|---:|---|---|---|---|---:|---:|---:|---:|---:|---:|")
    for name, kind, nmax, dither, cls, alpha in cases:
        img = make_image(w, h, cls, alpha)
        s, ln, info = study(L, kind, img, w, h, nmax, dither, a.starts, a.maxlen)
        ok = ln[ln > 0]
        med = int(np.median(ok)) if len(ok) else -1
        p90 = int(np.percentile(ok, 90)) if len(ok) else -1
        mx = int(ok.max()) if len(ok) else -1
        q = ("PriorityQueue" if info[1] else "ArrayDeque") + f" ({info[0]}, {info[2]})"
        print(f"| {name} | {nmax} | {'on' if dither else 'off'} | {cls}/{alpha} | {w}x{h} | {q} | {len(ln)} | {len(ok)} | "
              f"{med} | {p90} | {mx} | {int((ln < 0).sum())} |", flush=True)


if __name__ == "__main__":
    main()
