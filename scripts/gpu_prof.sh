#!/bin/bash
# ncu --set full on the conv and GroupNorm kernels of one N=44 step; exports small CSVs (reports stay on the box if big)
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT /tmp/ncu
CMD="python scripts/step_time.py --precision bf16 --batches 44 --iters 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 214 -c 40 -o /tmp/ncu/prof_conv $CMD > $OUT/ncu_conv.log 2>&1
echo ncu conv exit $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_ -s 244 -c 12 -o /tmp/ncu/prof_gn $CMD > $OUT/ncu_gn.log 2>&1
echo ncu gn exit $?
for n in prof_conv prof_gn; do
  ncu -i /tmp/ncu/$n.ncu-rep --page raw --csv > $OUT/$n.raw.csv 2>/dev/null
  ls -la /tmp/ncu/$n.ncu-rep
done
ls -la $OUT
