#!/bin/bash
# Evidence for profiles/: (1) ncu launch list of the bench command, (2) ncu --set full of the dominant kernel.
TAG=${1:-r1prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT /tmp/ncu
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/bench_under_ncu.log 2>&1
echo launches exit $?
CMD="python scripts/step_time.py --precision bf16 --batches 176 --iters 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 8 -c 6 -o /tmp/ncu/top $CMD > $OUT/ncu_top.log 2>&1
echo top exit $?
ncu -i /tmp/ncu/top.ncu-rep --page raw --csv > $OUT/top_halo.raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:gn_apply_kernel -s 2 -c 3 -o /tmp/ncu/gn $CMD > $OUT/ncu_gn.log 2>&1
ncu -i /tmp/ncu/gn.ncu-rep --page raw --csv > $OUT/gn_apply.raw.csv 2>/dev/null
ls -la $OUT
