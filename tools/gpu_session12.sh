#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/s12
timeout 800 python -m pytest tests/test_gpu_spec_dither.py tests/test_gpu_golden_big.py -x -q -k "spec or config3" > ${S}_pytest_spec.log 2>&1; echo "exit $?" >> ${S}_pytest_spec.log
tail -3 ${S}_pytest_spec.log
NQ_SPEC_REASONS=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > ${S}_bench1024.log 2> ${S}_bench1024.err; echo "exit $?" >> ${S}_bench1024.log
sort ${S}_bench1024.err | uniq -c | sort -rn | head -5
python - <<'PY'
import json
for f in ("gpurun_out/s12_bench1024.log",):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), round(d["ms_per_step"]), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"], d["kernels"]["k_spec_run"], d["roofline"]["frac"])
PY
