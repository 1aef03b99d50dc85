#!/bin/bash
# One gpurun call that takes the speculative dither from "CPU-verified" to "measured":
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_spec_session.sh'
# 1. its GPU parity tests, 2. per-launch device times on a 4K batch, 3. the bench with the path on,
# 4. ncu launch list and one full capture of k_spec_run (only after the plain runs exited 0).
set -u
mkdir -p gpurun_out

timeout 600 python -m pytest tests/test_gpu_spec_dither.py -x -q > gpurun_out/spec_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/spec_pytest.log

NQ_PROBE_NOORACLE=1 NQ_SPEC_TIMING=1 timeout 180 python tools/spec_gpu_probe.py 3840 2160 8192 1024 64 > gpurun_out/spec_probe_4k64.log 2>&1; echo "exit $?" >> gpurun_out/spec_probe_4k64.log
timeout 400 python bench.py --no-cpu --spec-dither 1 --steps 2 --warmup 1 > gpurun_out/bench_spec.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/bench_spec.log
if [ $rc -eq 0 ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_spec.csv \
    python bench.py --steps 1 --warmup 1 --batch 148 --no-e2e --no-cpu --spec-dither 1 > gpurun_out/ncu_launches_spec.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_spec_run -c 1 -o gpurun_out/prof_spec_run \
    python bench.py --steps 1 --warmup 1 --batch 64 --no-e2e --no-cpu --spec-dither 1 > gpurun_out/ncu_spec_run.log 2>&1
fi
tail -3 gpurun_out/spec_pytest.log gpurun_out/bench_spec.log
