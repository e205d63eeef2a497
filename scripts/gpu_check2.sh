#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest -q --tb=short -p no:cacheprovider "$@" > $OUT/$name.log 2>&1; echo "exit $?"; tail -n 12 $OUT/$name.log; }
run sampler tests/test_sampler_gpu.py -s
run gae tests/test_gae_gpu.py -s
run e2e tests/test_e2e_gpu.py -s
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python scripts/step_time.py --precision bf16 --batches 44,176 --iters 3 2>&1 | tee $OUT/step_bf16.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_n44.csv python scripts/step_time.py --precision bf16 --batches 44 --iters 1 > $OUT/ncu_stdout.log 2>&1
echo ncu exit $?
