#!/usr/bin/env python3
"""bench.py -- Mpixels/s of the quantizer hot path (convert = alpha scan + histogram + PNN merge +
palette + Gilbert dither) on batches of synthetic 4K images, one process per GPU, no collective.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                     the CPU arm: the oracle (C++ restatement of
                                                           the reference's Java core; no JVM exists here)
                                                           on all host cores, rank 0 only

A step is one pass of convert() over this rank's batch of images. `value` is measured with the
pixels already resident in HBM (nq_convert_batch_device), `e2e` through the host-buffer entry point
(nq_convert_batch: pinned host -> device -> host inside the timed region). Weak scaling: every rank
owns `--batch` images. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

KINDS = {"rgb": 0, "lab": 1}
CLASSES = {"smooth": 0, "noisy": 1, "rand": 2}
METRIC = "Mpixels/s 256-color PNN quantize+dither, 4K batch, 1/2/4/8 B200 vs JVM CPU"
# algorithmic bytes per pixel (SURVEY.md 8d / DESIGN.md): stage -> bytes
STAGE_BYTES = {"alpha_scan": 4, "histogram": 4, "find_nn_sweep": 0, "merge": 0, "dither_setup": 0, "dither": 8}
# dram__bytes_read.sum + dram__bytes_write.sum per pixel, from the ncu capture named in profiles/ (filled per round)
NCU_DRAM_BYTES_PER_PIXEL = {
    # profiles/r1_final_ncu_lab_8x1080p.md: k_dither_fifo read 236.1 MB + wrote 58.8 MB for 8 x 1920x1080 pixels
    # (8 B/px algorithmic + visiting-order table, RGB->Lab table sectors, candidate lists)
    ("lab", "dither"): (236.132352e6 + 58.803712e6) / (8 * 1920 * 1080),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("NQ_BENCH_BATCH", "592")),
                    help="images per GPU per step (592 = 4 merge CTAs on each of the 148 SMs)")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--kind", default="lab", choices=list(KINDS))
    ap.add_argument("--colors", type=int, default=256)
    ap.add_argument("--dither", type=int, default=1)
    ap.add_argument("--cls", default="noisy", choices=list(CLASSES))
    ap.add_argument("--spec-dither", type=int, default=int(os.environ.get("NQ_SPEC_DITHER", "0")),
                    help="1: speculative segment-parallel dither for the images that qualify (DESIGN.md 7.1; bit-identical results). "
                         "Off by default until its GPU tests have run on a B200")
    ap.add_argument("--spec-segment", type=int, default=8192)
    ap.add_argument("--spec-warmup", type=int, default=1024)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-workers", type=int, default=0, help="reference arm: worker processes (0 = all cores)")
    ap.add_argument("--cpu-width", type=int, default=1920, help="CPU legs: width of the bounded sample image")
    ap.add_argument("--cpu-height", type=int, default=1080, help="CPU legs: height of the bounded sample image")
    return ap.parse_args()


def workload_quantizer(a):
    return "PnnLABQuantizer" if a.kind == "lab" else "PnnQuantizer"


def workload_name(a):
    q = workload_quantizer(a)
    return f"{q} {a.colors} colors, dither {'on' if a.dither else 'off'}, batch of {a.batch} synthetic {a.cls} {a.width}x{a.height} ARGB images per GPU"


# --------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle on host cores
# --------------------------------------------------------------------------------------------------
def _cpu_one(job):
    kind, cls, w, h, colors, dither, seed, idx = job
    from oracle import pyoracle
    from nquant_android_b200.synth import make_image
    img = make_image(w, h, cls, "opaque", seed=0x5EED0000 + idx)
    t0 = time.perf_counter()
    pyoracle.convert(kind, img, w, h, colors, bool(dither), seed=seed + idx, trace=False)
    return time.perf_counter() - t0


def cpu_single(a):
    """one image, one thread: the like-for-like figure (the reference creates no threads). Bounded sample: one
    image of the workload's class and quantizer at --cpu-width x --cpu-height (a 4K CIELAB image costs the oracle
    about a minute; Mpixels/s is the unit, so the sample scales)."""
    from oracle import pyoracle
    pyoracle.build()
    w, h = min(a.cpu_width, a.width), min(a.cpu_height, a.height)
    dt = _cpu_one((KINDS[a.kind], a.cls, w, h, a.colors, a.dither, 0xC0FFEE, 0))
    return {"value": w * h / dt / 1e6, "unit": "Mpixels/s", "cores": 1, "kind": "port",
            "sample": f"1 image of the workload's class ({a.cls}, {workload_quantizer(a)}, {a.colors} colours, dither {a.dither}) at {w}x{h}, "
                      f"oracle/nq_oracle.cpp -O2, 1 thread, {dt:.1f} s; C++ restatement of the Java core (no JVM in the image)"}


def run_reference(a, rank):
    """--impl reference: the oracle (C++ restatement of the reference's Java core) on all host cores. Every step
    converts one bounded-sample image (--cpu-width x --cpu-height, the workload's class/quantizer/colours) per
    worker process."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pyoracle
    pyoracle.build()
    workers = a.cpu_workers or (os.cpu_count() or 1)
    kind = KINDS[a.kind]
    w, h = min(a.cpu_width, a.width), min(a.cpu_height, a.height)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for _ in range(a.warmup):   # untimed, on a small sample
            pool.map(_cpu_one, [(kind, a.cls, 256, 256, a.colors, a.dither, 0xC0FFEE, i) for i in range(workers)])
        t0 = time.perf_counter()
        for s in range(a.steps):
            pool.map(_cpu_one, [(kind, a.cls, w, h, a.colors, a.dither, 0xC0FFEE, s * workers + i) for i in range(workers)])
        dt = time.perf_counter() - t0
    px = a.steps * workers * w * h
    val = px / dt / 1e6
    sample = (f"{workers} images per step (one per worker process) of the workload's class at {w}x{h} "
              f"({a.cls}, {workload_quantizer(a)}, {a.colors} colours, dither {a.dither}); warm-up steps use 256x256 images")
    line = {"metric": METRIC, "value": val, "unit": "Mpixels/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a), "sample": sample},
            "cpu_baseline": {"value": val, "unit": "Mpixels/s", "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from nquant_android_b200.quantizer import Context

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_spec_dither(bool(a.spec_dither), a.spec_segment, a.spec_warmup)
    kind, cls = KINDS[a.kind], CLASSES[a.cls]
    npix = a.width * a.height
    n = a.batch
    din = torch.empty(n * npix, dtype=torch.int32, device="cuda")
    dout = torch.empty_like(din)
    ctx.synth_device(din.data_ptr(), n, a.width, a.height, cls, 0, 0x5EED0000 + rank * n)
    seeds = (np.arange(n, dtype=np.uint64) + np.uint64(0xC0FFEE + rank * n))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        ctx.convert_batch_ptr(kind, din.data_ptr(), dout.data_ptr(), n, a.width, a.height, a.colors, a.dither, seeds=seeds, device=True)

    for _ in range(a.warmup):
        step_device()
    ctx.stage_times(reset=True)
    l0 = ctx.kernel_launches()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step_device()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - l0
    stages = ctx.stage_times()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_px = world * n * npix * a.steps
    value = total_px / (ms_max / 1e3) / 1e6

    # ---- end to end: pinned host buffers through nq_convert_batch
    e2e = None
    if not a.no_e2e:
        # pinned host buffers for the whole batch (in + out); every local rank needs its own. If the host cannot hold
        # them next to everything else, the end-to-end leg runs on a smaller batch and says so.
        ne = n
        try:
            import psutil
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            avail = psutil.virtual_memory().available
            while ne > 1 and 2 * ne * npix * 4 * local_world > 0.6 * avail:
                ne //= 2
        except Exception:
            pass
        hin = torch.empty(ne * npix, dtype=torch.int32, pin_memory=True)
        hout = torch.empty(ne * npix, dtype=torch.int32, pin_memory=True)
        hin.copy_(din[:ne * npix])
        torch.cuda.synchronize()
        pal = np.zeros((ne, 256), dtype=np.uint32)
        plen = np.zeros(ne, dtype=np.int32)

        def step_host():
            ctx.convert_batch_ptr(kind, hin.data_ptr(), hout.data_ptr(), ne, a.width, a.height, a.colors, a.dither, seeds=seeds[:ne],
                                  device=False, palettes=pal, palette_lens=plen)

        step_host()   # allocates the staging buffers
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(a.steps):
            step_host()
        f1.record(stream)
        barrier()
        t2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item())
        same = bool(torch.equal(hout.cuda(), dout[:ne * npix]))
        e2e = {"value": world * ne * npix * a.steps / (e2e_ms / 1e3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": ne * npix * 4,
               "d2h_bytes_per_step": ne * npix * 4 + ne * 256 * 4 + ne * 4, "ms_per_step": e2e_ms / a.steps,
               "matches_device_path": same, "batch_images_per_gpu": ne}
        if ne != n:
            e2e["note"] = f"host memory holds pinned buffers for {ne} of the {n} images per GPU: end-to-end leg run on the smaller batch"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        per_stage = {}
        dom, dom_ms = None, -1.0          # dominant stage among those that move pixel bytes
        top, top_ms = None, -1.0          # dominant stage overall
        for name, (sms, ln) in stages.items():
            bytes_total = STAGE_BYTES[name] * n * npix * a.steps
            gbs = bytes_total / (sms / 1e3) / 1e9 if sms > 0 else 0.0
            per_stage[name] = {"ms_per_step": sms / a.steps, "launches_per_step": ln / a.steps, "share": sms / ms if ms > 0 else 0,
                               "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
            if sms > top_ms:
                top, top_ms = name, sms
            if STAGE_BYTES[name] > 0 and sms > dom_ms:
                dom, dom_ms = name, sms
        dom_launches = max(1, stages[dom][1])
        dom_bytes_per_launch = STAGE_BYTES[dom] * n * npix * a.steps / dom_launches
        achieved = dom_bytes_per_launch / (dom_ms / dom_launches / 1e3) / 1e9 if dom_ms > 0 else 0.0
        # DRAM bytes per pixel of the stage's kernels from one `ncu --set full` capture (profiles/), scaled to this launch
        tpp = NCU_DRAM_BYTES_PER_PIXEL.get((a.kind, dom))
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (tpp * n * npix * a.steps / dom_launches) if tpp else None, "peak_source": peak_src,
                    "note": "largest stage that moves pixel bytes; it is a distance-1 recurrence per image (dependent-issue latency bound, "
                            "DESIGN.md 4.1/4.2), so the HBM fraction is structurally tiny. The pixel passes' fractions are under `stages`.",
                    "dominant_by_time": {"stage": top, "share": top_ms / ms if ms > 0 else 0,
                                         "note": "find_nn sweep and merge loop work on histogram bins (<= 65 536 per image), not pixels: "
                                                 "0 algorithmic pixel bytes; issue/latency evidence in profiles/*_ncu_merge_sweep.md"},
                    "end_to_end_frac": (12.0 * total_px / (ms_max / 1e3) / 1e9) / (peak * world)}
        line = {"metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "global_batch_images": world * n, "pixels_per_step": world * n * npix,
                           "parallelism": f"one image shard per GPU x{world}, no collective",
                           "l2": "inputs larger than L2 (batch x 33 MB per image)",
                           "dither_path": ("speculative segments %d/%d (%s)" % (a.spec_segment, a.spec_warmup, ctx.spec_stats())) if a.spec_dither
                                          else "serial chain per image"},
                "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "stages": per_stage}
        if e2e:
            line["e2e"] = e2e
        if not a.no_cpu and world == 1:
            try:
                line["cpu_baseline"] = cpu_single(a)
            except Exception as ex:   # the oracle is a checker, never a dependency of the measured path
                line["cpu_baseline"] = {"value": None, "unit": "Mpixels/s", "cores": 1, "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    run_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
