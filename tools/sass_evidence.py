#!/usr/bin/env python3
"""profiles/r2_sass_evidence.md: per-kernel counts of the Blackwell-relevant SASS mnemonics in the built library
(cuobjdump -sass nquant_android_b200/libnquant_b200.so). Runs on the CPU container."""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "nquant_android_b200", "libnquant_b200.so")
COLS = [("UBLKCP", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"), ("UTMALDG", r"^UTMALDG"), ("FMNMX3", r"^FMNMX3"), ("FMNMX", r"^FMNMX(\.|$)"),
        ("FMUL", r"^FMUL"), ("FADD", r"^FADD"), ("FFMA", r"^FFMA"), ("DFMA", r"^DFMA"), ("DADD", r"^DADD"), ("DMUL", r"^DMUL"),
        ("LDG/LD .128", r"^LDG?\..*128"), ("BAR", r"^BAR"), ("REDUX/VOTE/SHFL", r"^(REDUX|VOTE|SHFL)")]
KEEP = ("k_spec_run", "k_spec_fill", "k_spec_resolve", "k_spec_pre", "k_spec_memo", "k_spec_scan", "k_radix_scatter", "k_dither_fifo",
        "k_dither_sorted", "k_merge_lab", "k_merge_rgb", "k_find_nn_lab", "k_find_nn_rgb", "k_lab_bin_sum", "k_alpha_scan", "k_bn_rgb")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            counts[cur]["n"] += 1
            for name, pat in COLS:
                if re.match(pat, op):
                    counts[cur][name] += 1
    dem = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    out = ["# r2: SASS evidence (`python tools/sass_evidence.py`: cuobjdump -sass nquant_android_b200/libnquant_b200.so, sm_100a)", "",
           "Instruction counts per kernel for the Blackwell-relevant mnemonics.", "",
           "* `UBLKCP.S.G` (PTX `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`) and `SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK.TRANS64.TRYWAIT`",
           "  (PTX `mbarrier.arrive.expect_tx`, `mbarrier.try_wait.parity`) are stage 6's optional record stream: each warp keeps a two-stage ring of 8 rows x 512 B",
           "  in shared memory, one elected lane issues the bulk copies of the next rows while the warp consumes the current ones (`WarpRing`, `run_span`,",
           "  nq_dither_spec.cuh). Measured neutral against the plain 128-bit loads (profiles/r2_bulk_ab.md), so `NQ_SPEC_BULK` is off by default.",
           "* `FMNMX3` is sm_100's three-input min/max (PTX `max.f32 d, a, b, c`), used by stage 6 for the running `maxErr` (GilbertCurve.java:192-201).",
           "* No `UTMALDG` (tensor-map TMA) and no `UTCMMA` (tcgen05): the path has no dense contraction and its tiles are 1-D runs of records, which the",
           "  non-tensor bulk copy covers.", "",
           "| kernel | instructions | " + " | ".join(c for c, _ in COLS) + " |", "|---|---:|" + "---:|" * len(COLS)]
    for (fn, c), d in zip(counts.items(), dem):
        if not any(k in fn for k in KEEP):
            continue
        name = re.sub(r"\(.*", "", d).replace("void ", "")
        out.append(f"| `{name}` | {c['n']} | " + " | ".join(str(c[col]) for col, _ in COLS) + " |")
    open(os.path.join(HERE, "..", "profiles", "r2_sass_evidence.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[-30:]))


if __name__ == "__main__":
    main()
