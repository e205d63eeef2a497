#!/bin/bash
# Subset of the GPU tests, then same-box A/B of the step latency: previous commit (build_tmp/libhsidm_head.so) vs working tree.
timeout 600 python -m pytest -q -x -p no:cacheprovider tests/test_unet_gpu.py -k "fused_attention_core or folded_attention" 2>&1 | tail -8
echo "== head";  STEP_LAT_N=5,176 HSIDM_AB_LIB=build_tmp/libhsidm_head.so timeout 300 python scripts/step_latency.py 2>&1 | tail -2
echo "== tree";  STEP_LAT_N=5,176 timeout 300 python scripts/step_latency.py 2>&1 | tail -2
echo "== tree two-kernel attention";  STEP_LAT_N=5,176 HSIDM_VARIANT=4096 timeout 300 python scripts/step_latency.py 2>&1 | tail -2
mkdir -p gpurun_out/r2g; python scripts/layer_prof.py --out gpurun_out/r2g/layer_prof_flash.csv > gpurun_out/r2g/layer_prof_flash.txt 2>&1; grep -E "total|attn|gemm|apply.T|k1 cin512.0 cout512 16x16" gpurun_out/r2g/layer_prof_flash.txt | head
