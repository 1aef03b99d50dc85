#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/s8
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -k "not 8192 and not full_size" > ${S}_pytest_parity.log 2>&1; echo "exit $?" >> ${S}_pytest_parity.log
tail -4 ${S}_pytest_parity.log
timeout 600 python -m pytest tests/test_gpu_golden_big.py -x -q -k "config3 or config1 or config0 or lab_256_on" > ${S}_pytest_golden.log 2>&1; echo "exit $?" >> ${S}_pytest_golden.log
tail -3 ${S}_pytest_golden.log
NQ_SPEC_REASONS=1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu > ${S}_bench1024.log 2> ${S}_bench1024.err; echo "exit $?" >> ${S}_bench1024.log
sort ${S}_bench1024.err | uniq -c | sort -rn | head -5
python - <<'PY'
import json
for f in ("gpurun_out/s8_bench1024.log",):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), round(d["ms_per_step"]), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"], d["kernels"]["k_spec_run"], d.get("e2e", {}).get("value"), d.get("golden"))
PY
