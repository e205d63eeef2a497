#!/bin/bash
mkdir -p gpurun_out/exp2
echo "== PDL on"; python scripts/step_latency.py 2>&1 | tee gpurun_out/exp2/pdl_on.txt
echo "== PDL off"; HSIDM_NO_PDL=1 python scripts/step_latency.py 2>&1 | tee gpurun_out/exp2/pdl_off.txt
echo "== tests"; timeout 900 python -m pytest -q --tb=short -p no:cacheprovider tests/test_sampler_gpu.py tests/test_unet_gpu.py tests/test_e2e_gpu.py tests/test_kernels_gpu.py -x 2>&1 | tail -8
