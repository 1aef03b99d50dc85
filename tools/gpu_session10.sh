#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/s10
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > ${S}_pool.log 2>&1; echo "exit $?" >> ${S}_pool.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --spec-segment 4096 > ${S}_seg4096.log 2>&1; echo "exit $?" >> ${S}_seg4096.log
NQ_SPEC_POOL_GB=48 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --spec-segment 4096 > ${S}_seg4096_pool48.log 2>&1; echo "exit $?" >> ${S}_seg4096_pool48.log
python - <<'PY'
import json
for f in ("gpurun_out/s10_pool.log", "gpurun_out/s10_seg4096.log", "gpurun_out/s10_seg4096_pool48.log"):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), round(d["ms_per_step"]), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"], d["kernels"]["k_spec_run"])
PY
