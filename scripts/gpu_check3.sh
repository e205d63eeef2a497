#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest -q --tb=short -p no:cacheprovider "$@" > $OUT/$name.log 2>&1; echo "exit $?"; tail -n 6 $OUT/$name.log; }
run kernels tests/test_kernels_gpu.py
run unet tests/test_unet_gpu.py
run sampler tests/test_sampler_gpu.py
run gae tests/test_gae_gpu.py
run e2e tests/test_e2e_gpu.py -s
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python scripts/step_time.py --precision bf16 --batches 5,44,176 --iters 3 2>&1 | tee $OUT/step_bf16.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 214 -c 40 -o $OUT/prof_conv python scripts/step_time.py --precision bf16 --batches 44 --iters 1 > $OUT/ncu_conv.log 2>&1
echo ncu conv exit $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_ -s 244 -c 12 -o $OUT/prof_gn python scripts/step_time.py --precision bf16 --batches 44 --iters 1 > $OUT/ncu_gn.log 2>&1
echo ncu gn exit $?
ls -la $OUT
