#!/usr/bin/env python3
"""Summarises an .ncu-rep (one `ncu --set full --import-source on` capture) into markdown: headline metrics per
kernel and the source lines with the most warp-stall samples. Usage: summarize_ncu.py report.ncu-rep "title" > out.md"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers, CTAs/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory, CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC per SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    print(f"# {title}\n")
    print(f"Source: `{rep}` (ncu --set full --clock-control none --import-source on). Numbers under the profiler are for\n"
          "shape, not for speed: bench.py's CUDA-event timings are the speed numbers.\n")
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    names = []
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").split("(")[0]
        if name.startswith("void "):
            name = name[5:]
        name = name.split("<")[0].strip()
        names.append(name)
        print(f"## `{name}`\n")
        print("| metric | value |")
        print("|---|---|")
        for k, label in KEYS:
            if k in d and d[k] not in ("", "nan", "-nan"):
                print(f"| {label} | {d[k]} {u.get(k, '')} |")
        print()
    for name in dict.fromkeys(names):
        out = ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + name.split("::")[-1]])
        fname, agg = None, {}
        for r in csv.reader(io.StringIO(out)):
            if not r:
                continue
            if r[0] == "File Path":
                fname = r[1].split("/")[-1]
                continue
            if r[0] in ("Function Name", "Line No"):
                continue
            if r[0] != "" and len(r) > 7 and r[2] == "-":
                try:
                    key = (fname, int(r[0]))
                    prev = agg.get(key, (0, 0, r[1]))
                    ex = int(float(r[7])) if r[7] not in ("", "nan", "-nan") else 0
                    agg[key] = (prev[0] + int(r[6]), prev[1] + ex, r[1])
                except ValueError:
                    pass
        tot_s = sum(v[0] for v in agg.values()) or 1
        tot_i = sum(v[1] for v in agg.values()) or 1
        print(f"### `{name}`: source lines by warp-stall samples ({tot_s} samples, {tot_i} warp instructions)\n")
        print("| file:line | samples | instructions | source |")
        print("|---|---:|---:|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
            src = v[2].strip().replace("|", "\\|")[:110]
            print(f"| {k[0]}:{k[1]} | {100 * v[0] / tot_s:.1f}% | {100 * v[1] / tot_i:.1f}% | `{src}` |")
        print()


if __name__ == "__main__":
    main()
