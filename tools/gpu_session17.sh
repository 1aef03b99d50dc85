#!/bin/bash
# k_merge_lab: the previous commit's library against today's default in the same session, then ncu --set full of the new kernel
set -u
mkdir -p gpurun_out
S=gpurun_out/s17
NQ_AB_LIB=ab/libnquant_old.so timeout 300 python tools/merge_mode_probe.py 3840 2160 592 0 > ${S}_merge_old.log 2>&1; echo "exit $?" >> ${S}_merge_old.log
tail -3 ${S}_merge_old.log
timeout 300 python tools/merge_mode_probe.py 3840 2160 592 18,0 > ${S}_merge_new.log 2>&1; echo "exit $?" >> ${S}_merge_new.log
tail -5 ${S}_merge_new.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_merge_lab -c 1 -o ${S}_prof_merge \
  python bench.py --steps 1 --warmup 0 --batch 592 --width 512 --height 512 --no-e2e --no-cpu > ${S}_ncu_merge.log 2>&1; echo "exit $?" >> ${S}_ncu_merge.log
tail -2 ${S}_ncu_merge.log
