#!/bin/bash
# k_merge_lab after the latency work (probe from the live-list record, heap prefetch, no second offsets pass): A/B of the
# instances and of the record prefetch, then parity of the candidates
set -u
mkdir -p gpurun_out
S=gpurun_out/s16
timeout 400 python tools/merge_mode_probe.py 3840 2160 592 0,2,18 > ${S}_merge_modes.log 2>&1; echo "exit $?" >> ${S}_merge_modes.log
cat ${S}_merge_modes.log | tail -8
for M in 18 2; do
  NQ_MERGE_MODE=$M timeout 300 python -m pytest tests/test_gpu_golden_big.py -x -q -k "config1 or config3 or config0 or q3_8192_lab" > ${S}_pytest_golden_m$M.log 2>&1; echo "exit $?" >> ${S}_pytest_golden_m$M.log
  tail -3 ${S}_pytest_golden_m$M.log
done
