#!/bin/bash
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT /tmp/ncu
CMD="python scripts/step_time.py --precision bf16 --batches 88 --iters 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 0 -c 2 -o /tmp/ncu/prof_h464 $CMD > $OUT/ncu_h464.log 2>&1
echo ncu exit $?
ncu -i /tmp/ncu/prof_h464.ncu-rep --page raw --csv > $OUT/prof_h464.raw.csv 2>/dev/null
ncu -i /tmp/ncu/prof_h464.ncu-rep --page source --csv --kernel-id :::2 > $OUT/prof_h464.source2.csv 2>/dev/null
ncu -i /tmp/ncu/prof_h464.ncu-rep --page details --kernel-id :::2 > $OUT/prof_h464.details2.txt 2>/dev/null
ls -la /tmp/ncu/ $OUT
