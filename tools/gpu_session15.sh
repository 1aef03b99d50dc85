#!/bin/bash
# k_merge_lab variants (NQ_MERGE_MODE): A/B on 592 x 4K, then parity of the candidates against the oracle fixtures
set -u
mkdir -p gpurun_out
S=gpurun_out/s15
timeout 400 python tools/merge_mode_probe.py 3840 2160 592 0,2,6,10,8 > ${S}_merge_modes.log 2>&1; echo "exit $?" >> ${S}_merge_modes.log
cat ${S}_merge_modes.log | tail -14
for M in 10 2; do
  NQ_MERGE_MODE=$M timeout 300 python -m pytest tests/test_gpu_golden_big.py -x -q -k "config1 or config3 or config0 or q3_8192_lab" > ${S}_pytest_golden_m$M.log 2>&1; echo "exit $?" >> ${S}_pytest_golden_m$M.log
  tail -3 ${S}_pytest_golden_m$M.log
done
NQ_MERGE_MODE=10 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q > ${S}_pytest_parity_m10.log 2>&1; echo "exit $?" >> ${S}_pytest_parity_m10.log
tail -3 ${S}_pytest_parity_m10.log
