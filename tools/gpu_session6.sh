#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/s6
timeout 800 python -m pytest tests/test_gpu_spec_dither.py -x -q > ${S}_pytest_spec.log 2>&1; echo "exit $?" >> ${S}_pytest_spec.log
tail -3 ${S}_pytest_spec.log
NQ_SPEC_REASONS=1 timeout 400 python bench.py --no-cpu --no-e2e --batch 592 --steps 2 --warmup 1 > ${S}_bench592.log 2> ${S}_bench592.err; echo "exit $?" >> ${S}_bench592.log
sort ${S}_bench592.err | uniq -c | sort -rn | head -10
NQ_SPEC_REASONS=1 timeout 900 python bench.py --steps 3 --warmup 2 > ${S}_bench1024.log 2> ${S}_bench1024.err; echo "exit $?" >> ${S}_bench1024.log
sort ${S}_bench1024.err | uniq -c | sort -rn | head -10
python - <<'PY'
import json
for f in ("gpurun_out/s6_bench592.log", "gpurun_out/s6_bench1024.log"):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), round(d["ms_per_step"]), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"], d["kernels"]["k_spec_run"], d.get("e2e", {}).get("value"), d.get("golden"), d.get("cpu_baseline", {}).get("value"))
PY
