#!/bin/bash
# Subset of the GPU tests, then same-box A/B of the step latency: previous commit (build_tmp/libhsidm_head.so) vs working tree.
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_kernels_gpu.py -k "not simt and not fp32" 2>&1 | tail -3
echo "== head";  STEP_LAT_N=5,176 HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -2
echo "== tree";  STEP_LAT_N=5,176 timeout 300 python scripts/step_latency.py 2>&1 | tail -2
echo "== head";  STEP_LAT_N=176 HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -1
echo "== tree";  STEP_LAT_N=176 timeout 300 python scripts/step_latency.py 2>&1 | tail -1
mkdir -p gpurun_out/r2g; python scripts/layer_prof.py --out gpurun_out/r2g/layer_prof.csv > gpurun_out/r2g/layer_prof.txt 2>&1; head -12 gpurun_out/r2g/layer_prof.txt
