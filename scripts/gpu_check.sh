#!/bin/bash
# Runs the GPU test suite in isolated processes (a poisoned CUDA context in one group must not hide the others)
# and drops logs under gpurun_out/.  Usage: scripts/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest -q --tb=short -p no:cacheprovider "$@" > $OUT/$name.log 2>&1; echo "exit $?"; tail -n 25 $OUT/$name.log; }
run k_simt tests/test_kernels_gpu.py -k "simt or groupnorm"
run k_tc tests/test_kernels_gpu.py -k "tc or dispatch" -s
run unet_fp32 tests/test_unet_gpu.py -k "not bf16" -s
run unet_bf16 tests/test_unet_gpu.py -k "bf16" -s
run sampler tests/test_sampler_gpu.py -s
run gae tests/test_gae_gpu.py -s
run e2e tests/test_e2e_gpu.py -s
run dropin tests/test_dropin_gpu.py -s
run scene tests/test_scene_gpu.py tests/test_prepost_gpu.py -s
