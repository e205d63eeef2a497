#!/bin/bash
# Subset of the GPU tests, then same-box A/B of the step latency: previous commit (build_tmp/libhsidm_head.so) vs working tree.
timeout 600 python -m pytest -q -x -p no:cacheprovider tests/test_kernels_gpu.py -k "tc_matches_torch and halo" 2>&1 | tail -15
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_unet_gpu.py tests/test_sampler_gpu.py -k "not fp32" 2>&1 | tail -6
echo "== head";  STEP_LAT_N=1,5,176 HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -3
echo "== tree";  STEP_LAT_N=1,5,176 timeout 300 python scripts/step_latency.py 2>&1 | tail -3
mkdir -p gpurun_out/r2g; python scripts/layer_prof.py --out gpurun_out/r2g/layer_prof_dual.csv > gpurun_out/r2g/layer_prof_dual.txt 2>&1; grep -E "total|8x8|apply" gpurun_out/r2g/layer_prof_dual.txt | head -14
