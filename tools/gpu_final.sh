#!/bin/bash
# round 2, final state: GPU tests, smoke, the default bench, its ncu launch list and one full capture of the dominant
# pixel-moving kernel. Everything lands in gpurun_out/ and is copied to profiles/ by hand afterwards.
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_final
timeout 1500 python -m pytest tests -q -m gpu > ${S}_pytest_gpu.log 2>&1; echo "exit $?" >> ${S}_pytest_gpu.log
tail -3 ${S}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > ${S}_smoke.log 2>&1; echo "exit $?" >> ${S}_smoke.log
tail -2 ${S}_smoke.log
timeout 1200 python bench.py > ${S}_bench.json 2> ${S}_bench.err; rc=$?; echo "bench exit $rc"
tail -c 600 ${S}_bench.json
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${S}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > ${S}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spec_run -c 1 -o ${S}_prof_spec_run \
    python bench.py --steps 1 --warmup 0 --batch 64 --no-e2e --no-cpu > ${S}_ncu_spec_run.log 2>&1; echo "ncu full exit $?"
fi
ls -la gpurun_out | grep r2_final
