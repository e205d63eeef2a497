#!/bin/bash
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT /tmp/ncu
CMD="python scripts/step_time.py --precision bf16 --batches 88 --iters 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 120 -c 14 -o /tmp/ncu/prof_halo $CMD > $OUT/ncu_halo.log 2>&1
echo ncu halo exit $?
ncu -i /tmp/ncu/prof_halo.ncu-rep --page raw --csv > $OUT/prof_halo.raw.csv 2>/dev/null
ncu -i /tmp/ncu/prof_halo.ncu-rep --page source --csv --kernel-id :::1 > $OUT/prof_halo.source1.csv 2>/dev/null
ls -la /tmp/ncu/ $OUT
