#!/bin/bash
mkdir -p gpurun_out/exp3
python bench.py --workload c1 --no-gpu-baseline > gpurun_out/exp3/bench_c1.json 2> gpurun_out/exp3/bench_c1.err; echo "c1 exit $?"; tail -c 300 gpurun_out/exp3/bench_c1.err
python bench.py --workload c3 --steps 40 --no-cpu --no-gpu-baseline > gpurun_out/exp3/bench_c3_n1.json 2> gpurun_out/exp3/bench_c3_n1.err; echo "c3 exit $?"; tail -c 300 gpurun_out/exp3/bench_c3_n1.err
python bench.py --steps 20 --gae-precision bf16 --no-cpu --no-gpu-baseline --no-e2e > gpurun_out/exp3/bench_gae_bf16.json 2> gpurun_out/exp3/bench_gae_bf16.err; echo "gae exit $?"
python - <<'PY'
import json
for f in ("bench_c1","bench_c3_n1","bench_gae_bf16"):
    try:
        d=json.load(open(f"gpurun_out/exp3/{f}.json"))
        print(f, {k:d.get(k) for k in ("value","ms_per_step","encode_ms","decode_ms","latents_per_step","tiles","non_sampling_ms","full_sampling")}, d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
