#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/s7
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > ${S}_pytest_gpu.log 2>&1; echo "exit $?" >> ${S}_pytest_gpu.log
tail -14 ${S}_pytest_gpu.log
NQ_PROBE_NOORACLE=1 NQ_SPEC_TIMING=1 timeout 240 python tools/spec_gpu_probe.py 3840 2160 0 1024 64 > ${S}_probe_4k64.log 2>&1; echo "exit $?" >> ${S}_probe_4k64.log
grep "pre \|scan \|resolve \|fill\|run " ${S}_probe_4k64.log | head -8
NQ_SPEC_REASONS=1 timeout 900 python bench.py --steps 3 --warmup 2 > ${S}_bench1024.log 2> ${S}_bench1024.err; echo "exit $?" >> ${S}_bench1024.log
sort ${S}_bench1024.err | uniq -c | sort -rn | head -5
python - <<'PY'
import json
for f in ("gpurun_out/s7_bench1024.log",):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), round(d["ms_per_step"]), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"], d["kernels"]["k_spec_run"], d.get("e2e", {}).get("value"), d.get("golden"), d.get("cpu_baseline", {}).get("value"))
PY
