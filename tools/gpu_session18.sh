#!/bin/bash
# end of round 2: one run of every BASELINE.json config (configs[4]: the 256-colour dither-on rows), then the whole GPU suite
set -u
mkdir -p gpurun_out
S=gpurun_out/s18
timeout 330 python tools/config_table.py --big256 > ${S}_configs.md 2> ${S}_configs.err; echo "config table exit $?"
cat ${S}_configs.md
timeout 760 python -m pytest tests -q -m gpu > ${S}_pytest_gpu.log 2>&1; echo "exit $?" >> ${S}_pytest_gpu.log
tail -4 ${S}_pytest_gpu.log
