#!/bin/bash
# round 2, session 4: sequential chains in the speculative dither, alpha-bitmask variant of stage 6, merge warp rotation A/B
set -u
mkdir -p gpurun_out
S=gpurun_out/s4
timeout 700 python -m pytest tests/test_gpu_spec_dither.py -x -q > ${S}_pytest_spec.log 2>&1; echo "exit $?" >> ${S}_pytest_spec.log
tail -3 ${S}_pytest_spec.log
NQ_PROBE_NOORACLE=1 NQ_SPEC_TIMING=1 timeout 240 python tools/spec_gpu_probe.py 3840 2160 0 1024 64 > ${S}_probe_4k64.log 2>&1; echo "exit $?" >> ${S}_probe_4k64.log
grep -c "round" ${S}_probe_4k64.log
timeout 400 python bench.py --no-cpu --no-e2e --batch 592 --steps 2 --warmup 1 > ${S}_bench592.log 2>&1; echo "exit $?" >> ${S}_bench592.log
tail -2 ${S}_bench592.log | cut -c1-1000
NQ_MERGE_ROT=0 timeout 300 python bench.py --no-cpu --no-e2e --batch 592 --width 512 --height 512 --steps 2 --warmup 1 > ${S}_rot0.log 2>&1; echo "exit $?" >> ${S}_rot0.log
NQ_MERGE_ROT=1 timeout 300 python bench.py --no-cpu --no-e2e --batch 592 --width 512 --height 512 --steps 2 --warmup 1 > ${S}_rot1.log 2>&1; echo "exit $?" >> ${S}_rot1.log
python - <<'PY'
import json
for f in ("gpurun_out/s4_rot0.log", "gpurun_out/s4_rot1.log", "gpurun_out/s4_bench592.log"):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, round(d["value"], 1), {k: round(v["ms_per_step"], 1) for k, v in d["stages"].items()}, d["config"]["dither_path"])
PY
timeout 700 python bench.py --no-cpu --steps 2 --warmup 1 > ${S}_bench1024.log 2>&1; echo "exit $?" >> ${S}_bench1024.log
tail -2 ${S}_bench1024.log | cut -c1-1000
