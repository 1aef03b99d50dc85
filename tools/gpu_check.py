#!/usr/bin/env python3
"""Developer parity probe: runs the CUDA path and the oracle on the same synthetic image and reports
stage by stage where they first differ. Usage: gpu_check.py [kind cls alpha W H K dither]..."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from oracle import pyoracle as o
from nquant_android_b200.synth import make_image
from nquant_android_b200.quantizer import Context

KEYS = [("has_semi_transparency", "hasSemiTransparency"), ("transparent_pixel_index", "transparentPixelIndex"),
        ("maxbins", "maxbins"), ("quan_rt", "quan_rt"), ("weight", "weight"), ("pr", "PR"), ("pg", "PG"), ("pa", "PA"),
        ("g_margin", "margin"), ("g_thresold", "thresold"), ("g_dither_max_q", "DITHER_MAX"), ("g_dither_max", "ditherMax"),
        ("g_sorted", "sortedByYDiff"), ("g_has_alpha", "hasAlpha"), ("g_beta", "beta"), ("rng_draws", "rng_draws"),
        ("merges", "n_merges"), ("ratio_merge", "ratio_merge"), ("is_nano", "isNano"), ("texicab", "texicab")]


def check(ctx, kind, cls, alpha, W, H, K, dither, seed=0xC0FFEE, verbose=True):
    img = make_image(W, H, cls, alpha)
    t0 = time.time()
    ref = o.convert(kind, img, W, H, K, dither, seed=seed)
    t1 = time.time()
    ctx.set_debug(True)
    out, pal, plen, ha = ctx.convert_batch(kind, img[None, :], W, H, K, dither, seeds=[seed])
    t2 = time.time()
    info = ctx.image_info(0)
    tag = f"kind={kind} {cls}/{alpha} {W}x{H} K={K} dither={int(dither)}"
    ok = True
    msgs = []
    for gk, rk in KEYS:
        gv, rv = info[gk], ref.scalars[rk]
        if rk in ("isNano", "texicab", "ratio_merge") and kind == 0 and rk != "ratio_merge":
            continue
        if rk == "isNano" or (K <= 2 and rk in ("maxbins", "quan_rt", "n_merges", "ratio_merge", "PR", "PG", "PA")):
            continue
        if isinstance(rv, float) or isinstance(gv, float):
            same = np.float64(gv) == np.float64(rv) or (rk == "beta" and np.float32(gv) == np.float32(rv))
        else:
            same = int(gv) == int(rv)
        if not same:
            ok = False
            msgs.append(f"scalar {gk}: gpu={gv} ref={rv}")
    if K > 2 and len(ref.bins):
        bins, ierr, inn = ctx.debug_bins(0)
        if bins.shape != ref.bins.shape:
            ok = False; msgs.append(f"bins shape {bins.shape} vs {ref.bins.shape}")
        else:
            nb = int((bins != ref.bins).any(axis=1).sum())
            if nb:
                ok = False
                j = int(np.argmax((bins != ref.bins).any(axis=1)))
                msgs.append(f"bins differ in {nb} rows; first {j}: gpu={bins[j]} ref={ref.bins[j]}")
            ne = int((ierr.view(np.uint32) != ref.init_err.view(np.uint32)).sum()); nn_ = int((inn != ref.init_nn).sum())
            if ne or nn_:
                ok = False
                j = int(np.argmax((ierr.view(np.uint32) != ref.init_err.view(np.uint32)) | (inn != ref.init_nn)))
                msgs.append(f"init find_nn differs: err {ne}, nn {nn_}; first {j}: gpu=({ierr[j]},{inn[j]}) ref=({ref.init_err[j]},{ref.init_nn[j]})")
        mg = ctx.debug_merges(0)
        if mg.shape != ref.merges.shape:
            ok = False; msgs.append(f"merge count {mg.shape} vs {ref.merges.shape}")
        else:
            d = (mg != ref.merges).any(axis=1)
            if d.any():
                ok = False
                j = int(np.argmax(d))
                msgs.append(f"merge sequence differs from step {j} of {len(mg)}: gpu={mg[j]} ref={ref.merges[j]}")
    gp = pal[0, :plen[0]]
    if len(gp) != len(ref.palette) or (gp != ref.palette).any():
        ok = False
        nd = int((gp != ref.palette).sum()) if len(gp) == len(ref.palette) else -1
        msgs.append(f"palette differs ({nd} entries; len {len(gp)} vs {len(ref.palette)})")
    if len(ref.saliencies) and info["g_use_saliency"]:
        sal = ctx.debug_saliencies(W * H, 0)
        ns = int((sal.view(np.uint32) != ref.saliencies.view(np.uint32)).sum())
        if ns:
            ok = False; msgs.append(f"saliency differs at {ns} pixels")
    npx = int((out[0] != ref.out).sum())
    if npx:
        ok = False
        order = o.gilbert_order(W, H)
        first = int(np.argmax(out[0][order] != ref.out[order]))
        msgs.append(f"output differs at {npx}/{W*H} pixels; first along the curve at step {first} (bidx {order[first]}): gpu={out[0][order[first]]:08x} ref={ref.out[order[first]]:08x}")
    print(("PASS " if ok else "FAIL ") + tag + f"  oracle {t1-t0:.2f}s gpu {t2-t1:.2f}s bins={info['maxbins']} rescans={info['rescans']} pairs={info['pair_tests']} draws={info['rng_draws']}", flush=True)
    if verbose:
        for m in msgs:
            print("     " + m, flush=True)
    return ok


if __name__ == "__main__":
    ctx = Context(0)
    cases = []
    args = sys.argv[1:]
    if args:
        for i in range(0, len(args), 7):
            k, cls, al, W, H, K, d = args[i:i + 7]
            cases.append((int(k), cls, al, int(W), int(H), int(K), bool(int(d))))
    else:
        for kind in (0, 1):
            for cls in ("smooth", "noisy"):
                cases.append((kind, cls, "opaque", 128, 96, 256, True))
        cases += [(0, "noisy", "opaque", 128, 96, 16, True), (0, "noisy", "semi", 128, 96, 16, True),
                  (0, "noisy", "transparent", 128, 96, 64, True), (0, "noisy", "opaque", 128, 96, 256, False),
                  (0, "smooth", "opaque", 128, 96, 2, True), (1, "noisy", "semi", 128, 96, 16, True),
                  (1, "noisy", "opaque", 128, 96, 64, True), (1, "smooth", "opaque", 128, 96, 16, True),
                  (1, "smooth", "opaque", 128, 96, 4, True), (1, "smooth", "opaque", 128, 96, 2, True),
                  (0, "rand", "opaque", 256, 256, 256, True), (1, "rand", "opaque", 256, 256, 256, True)]
    bad = 0
    for c in cases:
        try:
            bad += not check(ctx, *c)
        except Exception as e:
            bad += 1
            print("ERROR", c, repr(e), flush=True)
    print(f"{len(cases) - bad}/{len(cases)} cases pass")
    sys.exit(1 if bad else 0)
