#!/bin/bash
# round 2, session 2: new orchestration + speculative dither default. Every step under its own timeout.
set -u
mkdir -p gpurun_out
S=gpurun_out/s2
timeout 600 python -m pytest tests/test_gpu_spec_dither.py -x -q > ${S}_pytest_spec.log 2>&1; echo "exit $?" >> ${S}_pytest_spec.log
tail -3 ${S}_pytest_spec.log
NQ_PROBE_NOORACLE=1 NQ_SPEC_TIMING=1 timeout 240 python tools/spec_gpu_probe.py 3840 2160 0 1024 64 > ${S}_probe_4k64.log 2>&1; echo "exit $?" >> ${S}_probe_4k64.log
tail -4 ${S}_probe_4k64.log | cut -c1-600
timeout 400 python bench.py --no-cpu --batch 592 --steps 2 --warmup 1 > ${S}_bench592.log 2>&1; echo "exit $?" >> ${S}_bench592.log
tail -2 ${S}_bench592.log | cut -c1-1500
timeout 600 python bench.py --no-cpu --steps 2 --warmup 1 > ${S}_bench1024.log 2>&1; echo "exit $?" >> ${S}_bench1024.log
tail -2 ${S}_bench1024.log | cut -c1-1500
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_spec_dither.py --durations=15 > ${S}_pytest_gpu.log 2>&1; echo "exit $?" >> ${S}_pytest_gpu.log
tail -25 ${S}_pytest_gpu.log
