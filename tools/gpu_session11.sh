#!/bin/bash
# two GPUs: the torchrun path of bench.py with the strong-scaling leg (a reduced batch keeps it short)
set -u
mkdir -p gpurun_out
S=gpurun_out/s11
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 --batch 256 --strong-batch 256 > ${S}_bench2gpu.log 2> ${S}_bench2gpu.err; echo "exit $?" >> ${S}_bench2gpu.log
tail -3 ${S}_bench2gpu.log | cut -c1-3000
tail -5 ${S}_bench2gpu.err
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "multi" > ${S}_pytest_multi.log 2>&1; echo "exit $?" >> ${S}_pytest_multi.log
tail -3 ${S}_pytest_multi.log
