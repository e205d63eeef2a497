#!/bin/bash
# Same-box A/B of two builds of the library: build_tmp/libhsidm_head.so (the previous commit) against the working tree.
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_unet_gpu.py -k "folded or fused_input or bf16" 2>&1 | tail -4
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_sampler_gpu.py tests/test_kernels_gpu.py -k "not simt" 2>&1 | tail -3
echo "== head";  STEP_LAT_N=1,5,11,176 HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -4
echo "== tree";  STEP_LAT_N=1,5,11,176 timeout 300 python scripts/step_latency.py 2>&1 | tail -4
echo "== tree, finalize launches"; STEP_LAT_N=5,176 HSIDM_VARIANT=256 timeout 300 python scripts/step_latency.py 2>&1 | tail -2
mkdir -p gpurun_out/gnfold; python scripts/layer_prof.py --out gpurun_out/gnfold/layer_prof3.csv > gpurun_out/gnfold/layer_prof3.txt 2>&1; head -16 gpurun_out/gnfold/layer_prof3.txt
