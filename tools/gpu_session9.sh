#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/s9
timeout 300 python bench.py --no-cpu --no-e2e --batch 592 --width 512 --height 512 --steps 2 --warmup 1 > ${S}_merge512.log 2>&1; echo "exit $?" >> ${S}_merge512.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu > ${S}_bench1024.log 2> ${S}_bench1024.err; echo "exit $?" >> ${S}_bench1024.log
python - <<'PY'
import json
for f in ("gpurun_out/s9_merge512.log", "gpurun_out/s9_bench1024.log"):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), round(d["ms_per_step"]), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"], d["kernels"]["k_spec_run"], d.get("e2e", {}))
PY
