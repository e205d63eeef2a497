#!/bin/bash
# Template for a same-box A/B: a subset of the GPU tests, then the step latency of the previous build
# (build_tmp/libhsidm_head.so, copied there before rebuilding) against the working tree.
timeout 600 python -m pytest -q -x -p no:cacheprovider tests/test_unet_gpu.py -k "bf16 and reference" 2>&1 | tail -3
for i in 1 2; do
echo "== head";  STEP_LAT_N=176 HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -1
echo "== tree";  STEP_LAT_N=176 timeout 300 python scripts/step_latency.py 2>&1 | tail -1
done
mkdir -p gpurun_out/ab; python scripts/layer_prof.py --out gpurun_out/ab/layer_prof.csv > gpurun_out/ab/layer_prof.txt 2>&1; head -12 gpurun_out/ab/layer_prof.txt
