#!/usr/bin/env python3
"""Stage-time probe: device-resident batch through nq_convert_batch_device.
Usage: perf_probe.py kind cls W H K dither batch [repeat]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from nquant_android_b200.quantizer import Context

def main():
    a = sys.argv[1:]
    kind, cls, W, H, K, dither, batch = int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), int(a[6])
    rep = int(a[7]) if len(a) > 7 else 1
    ctx = Context(0)
    npix = W * H
    din = torch.empty(batch * npix, dtype=torch.int32, device="cuda")
    dout = torch.empty_like(din)
    ctx.synth_device(din.data_ptr(), batch, W, H, cls, 0, 0x5EED0000)
    seeds = np.arange(batch, dtype=np.uint64) + 0xC0FFEE
    for r in range(rep):
        ctx.stage_times(reset=True)
        torch.cuda.synchronize()
        t0 = time.time()
        ctx.convert_batch_ptr(kind, din.data_ptr(), dout.data_ptr(), batch, W, H, K, dither, seeds=seeds, device=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = ctx.stage_times()
        info = ctx.image_info(0)
        print(f"kind={kind} cls={cls} {W}x{H} K={K} dither={dither} batch={batch}: {dt:.3f}s  {batch*npix/dt/1e6:.2f} Mpix/s  "
              f"bins={info['maxbins']} rescans={info['rescans']} pairs={info['pair_tests']} full_evals={info['full_evals']}", flush=True)
        print("   " + "  ".join(f"{k}={v[0]:.1f}ms" for k, v in st.items()), flush=True)
        dc = info["dither_cycles"]
        print(f"   dither Mcycles: consumer={dc[0] / 1e6:.0f} (waiting {dc[1] / 1e6:.0f})  producer waiting={dc[2] / 1e6:.0f}", flush=True)
        if kind == 1:
            mc = info["merge_cycles"]
            names = ["heap", "first32", "blocktest", "screen", "full", "merge"]
            print("   merge Mcycles: " + "  ".join(f"{n}={c / 1e6:.0f}" for n, c in zip(names, mc)) +
                  f"  live_blocks={info['live_blocks']} screened={info['screened']}", flush=True)

main()
