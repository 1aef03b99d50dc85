#!/bin/bash
# round 2, session 3: faster error-dependent lookups in stage 6, all mispredictions corrected per round, BlueNoise per-pixel pass
set -u
mkdir -p gpurun_out
S=gpurun_out/s3
timeout 600 python -m pytest tests/test_gpu_spec_dither.py -x -q > ${S}_pytest_spec.log 2>&1; echo "exit $?" >> ${S}_pytest_spec.log
tail -3 ${S}_pytest_spec.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "False or demo" > ${S}_pytest_bn.log 2>&1; echo "exit $?" >> ${S}_pytest_bn.log
tail -3 ${S}_pytest_bn.log
NQ_PROBE_NOORACLE=1 NQ_SPEC_TIMING=1 timeout 240 python tools/spec_gpu_probe.py 3840 2160 0 1024 64 > ${S}_probe_4k64.log 2>&1; echo "exit $?" >> ${S}_probe_4k64.log
grep -c round ${S}_probe_4k64.log
timeout 400 python bench.py --no-cpu --no-e2e --batch 592 --steps 2 --warmup 1 > ${S}_bench592.log 2>&1; echo "exit $?" >> ${S}_bench592.log
tail -2 ${S}_bench592.log | cut -c1-1200
timeout 700 python bench.py --no-cpu --steps 2 --warmup 1 > ${S}_bench1024.log 2>&1; echo "exit $?" >> ${S}_bench1024.log
tail -2 ${S}_bench1024.log | cut -c1-1200
# merge loop profile: 592 images of 512x512 (the merge loop works on bins, ~27 k per image whatever the image size)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_merge_lab -c 1 -o ${S}_prof_merge \
  python bench.py --steps 1 --warmup 0 --batch 592 --width 512 --height 512 --no-e2e --no-cpu > ${S}_ncu_merge.log 2>&1; echo "exit $?" >> ${S}_ncu_merge.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spec_run -c 1 -o ${S}_prof_spec_run \
  python bench.py --steps 1 --warmup 0 --batch 64 --no-e2e --no-cpu > ${S}_ncu_spec_run.log 2>&1; echo "exit $?" >> ${S}_ncu_spec_run.log
ls -la gpurun_out | grep s3_
