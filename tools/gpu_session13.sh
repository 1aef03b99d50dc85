#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/s13
timeout 400 python bench.py --no-cpu --no-e2e --batch 592 --steps 2 --warmup 1 > ${S}_bench592.log 2>&1; echo "exit $?" >> ${S}_bench592.log
NQ_PROBE_NOORACLE=1 NQ_SPEC_TIMING=1 timeout 300 python tools/spec_gpu_probe.py 3840 2160 0 1024 160 > ${S}_probe_4k160.log 2>&1; echo "exit $?" >> ${S}_probe_4k160.log
python - <<'PY'
import json
for f in ("gpurun_out/s13_bench592.log",):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), round(d["ms_per_step"]), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"], d["kernels"]["k_spec_run"], d["roofline"]["frac"])
PY
grep "round\|run  " gpurun_out/s13_probe_4k160.log | head -60
