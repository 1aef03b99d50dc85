#!/usr/bin/env python3
"""Runs BASELINE.json's configs (and the three synthetic classes) once each on cuda:0 and prints a markdown table:
device-resident wall time of nq_convert_batch_device, Mpixels/s and the per-stage device times.
Usage: config_table.py [--big | --big256] > profiles/rNN_configs.md     (--big adds the 8192x8192 sweep of configs[4],
--big256 only its 256-colour dither-on rows)"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from nquant_android_b200.quantizer import Context

CLS = {"smooth": 0, "noisy": 1, "rand": 2}
ALPHA = {"opaque": 0, "transparent": 1, "semi": 2}


def _oracle_seconds():
    """oracle_seconds of the frozen cases (tests/golden/oracle_big_cases.json): the CPU time of ONE image on one core of the
    build container (measured while 4-5 oracle jobs shared its 8 cores, so a little high)."""
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "oracle_big_cases.json")
    out = {}
    try:
        for c in json.load(open(path)):
            out[(c["kind"], c["cls"], c["alpha"], c["w"], c["h"], c["k"], c["dither"])] = c.get("oracle_seconds")
    except Exception:
        pass
    return out


ORACLE_S = _oracle_seconds()


def run(ctx, label, kind, cls, alpha, W, H, K, dither, batch, reps=2):
    npix = W * H
    din = torch.empty(batch * npix, dtype=torch.int32, device="cuda")
    dout = torch.empty_like(din)
    ctx.synth_device(din.data_ptr(), batch, W, H, CLS[cls], ALPHA[alpha], 0x5EED0000)
    seeds = np.arange(batch, dtype=np.uint64) + 0xC0FFEE
    best = None
    for rep in range(reps):       # first run warms the workspace, the Gilbert order and the tables
        ctx.stage_times(reset=True)
        torch.cuda.synchronize()
        t0 = time.time()
        ctx.convert_batch_ptr(kind, din.data_ptr(), dout.data_ptr(), batch, W, H, K, dither, seeds=seeds, device=True)
        torch.cuda.synchronize()
        best = (time.time() - t0, ctx.stage_times(), ctx.image_info(0))
    dt, st, info = best
    q = "PnnLABQuantizer" if kind else "PnnQuantizer"
    stages = " / ".join(f"{v[0]:.0f}" for v in st.values())
    osec = ORACLE_S.get((kind, cls, alpha, W, H, K, int(bool(dither))))
    cpu = f"{npix / osec / 1e6:.2f}" if osec else "-"
    ratio = f"{(batch * npix / dt) / (npix / osec):.0f}x" if osec else "-"
    print(f"| {label} | {q} | {K} | {'on' if dither else 'off'} | {cls}/{alpha} | {W}x{H} | {batch} | {info['maxbins']} | "
          f"{dt * 1e3:.0f} | {batch * npix / dt / 1e6:.1f} | {cpu} | {ratio} | {stages} |", flush=True)
    del din, dout
    torch.cuda.empty_cache()


def main():
    big = "--big" in sys.argv
    big256 = "--big256" in sys.argv     # only the 256-colour, dither-on rows of configs[4] (one repetition: no warm run)
    ctx = Context(0)
    print("| config | quantizer | colours | dither | class/alpha | size | images | bins | ms | Mpixels/s | oracle Mpixels/s (1 core) | GPU / oracle | stage ms (scan / histogram / sweep / merge / setup / dither) |")
    print("|---|---|---:|---|---|---|---:|---:|---:|---:|---:|---:|---|")
    run(ctx, "configs[0]", 0, "noisy", "opaque", 512, 512, 256, 1, 1)
    run(ctx, "configs[1]", 1, "noisy", "opaque", 1920, 1080, 256, 1, 1)
    run(ctx, "configs[1] x592", 1, "noisy", "opaque", 1920, 1080, 256, 1, 592)
    run(ctx, "configs[2]", 0, "noisy", "semi", 3840, 2160, 16, 1, 1)
    run(ctx, "configs[3], all 1024 on one GPU", 1, "noisy", "opaque", 3840, 2160, 256, 1, 1024)
    for cls in ("smooth", "rand"):
        run(ctx, f"class {cls}", 1, cls, "opaque", 3840, 2160, 256, 1, 148)
        run(ctx, f"class {cls}", 0, cls, "opaque", 3840, 2160, 256, 1, 148)
    if big256:
        for kind in (1, 0):
            run(ctx, "configs[4]", kind, "noisy", "opaque", 8192, 8192, 256, 1, 1, reps=1)
    if big:
        for kind in (1, 0):
            for K in (256, 64, 16, 2):
                run(ctx, "configs[4]", kind, "noisy", "opaque", 8192, 8192, K, 1, 1)
            for K in (256, 64):
                run(ctx, "configs[4]", kind, "noisy", "opaque", 8192, 8192, K, 0, 1)


main()
