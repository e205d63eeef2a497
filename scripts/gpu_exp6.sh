#!/bin/bash
mkdir -p gpurun_out/exp6
echo "== base"; python scripts/step_latency.py 2>&1 | tail -2
echo "== no finalize"; HSIDM_SKIP_FINALIZE=1 python scripts/step_latency.py 2>&1 | tail -2
echo "== PDL everywhere"; HSIDM_PDL_MAX_PIXELS=100000000 python scripts/step_latency.py 2>&1 | tail -1
timeout 600 python -m pytest -q --tb=short -p no:cacheprovider tests/test_unet_gpu.py -k c4 -s 2>&1 | tail -6
python bench.py --workload c4 --steps 20 > gpurun_out/exp6/bench_c4.json 2> gpurun_out/exp6/bench_c4.err; echo "c4 exit $?"; tail -c 300 gpurun_out/exp6/bench_c4.err; cut -c1-700 gpurun_out/exp6/bench_c4.json
