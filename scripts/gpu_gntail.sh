#!/bin/bash
# GroupNorm statistics folded by the consumers (no gn_finalize launches) + folded attention projections: the tests that
# exercise them, then A/B timing against variant 256 (finalize launches) and 512 (materialised q, k, v).
OUT=gpurun_out/gnfold
mkdir -p $OUT
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest -q --tb=short -p no:cacheprovider "$@" > $OUT/$name.log 2>&1; echo "exit $?"; tail -n 14 $OUT/$name.log; }
run fold tests/test_unet_gpu.py -k "folded or fused_input" -s
run unet_bf16 tests/test_unet_gpu.py -k "bf16 and not folded and not fused_input" -s
run sampler tests/test_sampler_gpu.py -s
run kernels tests/test_kernels_gpu.py -k "tc or dispatch"
echo "=== step latency (default)"; timeout 600 python scripts/step_latency.py 2>&1 | tail -4
echo "=== step latency (finalize launches)"; HSIDM_VARIANT=256 timeout 600 python scripts/step_latency.py 2>&1 | tail -4
echo "=== step latency (materialised qkv)"; HSIDM_VARIANT=512 timeout 600 python scripts/step_latency.py 2>&1 | tail -4
echo "=== layer profile"; timeout 300 python scripts/layer_prof.py --out $OUT/layer_prof.csv > $OUT/layer_prof.txt 2>&1; head -30 $OUT/layer_prof.txt
