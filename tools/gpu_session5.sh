#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/s5
NQ_SPEC_REASONS=1 timeout 400 python bench.py --no-cpu --no-e2e --batch 592 --steps 2 --warmup 1 > ${S}_bench592.log 2> ${S}_bench592.err; echo "exit $?" >> ${S}_bench592.log
sort ${S}_bench592.err | uniq -c | sort -rn | head -20
python - <<'PY'
import json
for f in ("gpurun_out/s5_bench592.log",):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"], d["kernels"]["k_spec_run"])
PY
NQ_PROBE_NOORACLE=1 NQ_SPEC_TIMING=1 timeout 240 python tools/spec_gpu_probe.py 3840 2160 0 1024 64 > ${S}_probe_4k64.log 2>&1; echo "exit $?" >> ${S}_probe_4k64.log
grep "run \|round" ${S}_probe_4k64.log | head -30
