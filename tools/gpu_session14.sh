#!/bin/bash
# stage-6 record stream through cp.async.bulk + mbarrier (NQ_SPEC_BULK=1): parity and an A/B of the first launch
set -u
mkdir -p gpurun_out
S=gpurun_out/s14
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "stage_hook" > ${S}_pytest_hook.log 2>&1; echo "exit $?" >> ${S}_pytest_hook.log
tail -2 ${S}_pytest_hook.log
NQ_SPEC_BULK=1 timeout 420 python -m pytest tests/test_gpu_spec_dither.py -x -q > ${S}_pytest_spec_bulk.log 2>&1; echo "exit $?" >> ${S}_pytest_spec_bulk.log
tail -3 ${S}_pytest_spec_bulk.log
for B in 0 1; do
  NQ_SPEC_BULK=$B NQ_PROBE_NOORACLE=1 NQ_SPEC_TIMING=1 timeout 200 python tools/spec_gpu_probe.py 3840 2160 0 1024 64 > ${S}_probe_bulk$B.log 2>&1; echo "exit $?" >> ${S}_probe_bulk$B.log
  echo "bulk=$B"; grep "run  " ${S}_probe_bulk$B.log | head -4; grep "spec output\|exit" ${S}_probe_bulk$B.log | head -3
done
NQ_SPEC_BULK=1 timeout 500 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > ${S}_bench_bulk1.log 2>&1; echo "exit $?" >> ${S}_bench_bulk1.log
python - <<'PY'
import json
for f in ("gpurun_out/s14_bench_bulk1.log",):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), round(d["ms_per_step"]), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["kernels"]["k_spec_run"], d.get("golden"))
PY
