#!/bin/bash
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_kernels_gpu.py -k "not simt and not fp32" 2>&1 | tail -3
echo "== tree";  STEP_LAT_N=1,2,5,11,22,44 timeout 300 python scripts/step_latency.py 2>&1 | tail -6
echo "== tree, variant 1024";  STEP_LAT_N=1,2,5,11,22,44 HSIDM_VARIANT=1024 timeout 300 python scripts/step_latency.py 2>&1 | tail -6
