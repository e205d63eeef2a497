#!/bin/bash
# round 2, first GPU pass: whole suite, a short bench line, per-layer profile
bash scripts/gpu_check.sh r2a
python bench.py --steps 60 --warmup 3 > gpurun_out/r2a/bench_k60.json 2> gpurun_out/r2a/bench_k60.err; echo "bench exit $?"; tail -c 600 gpurun_out/r2a/bench_k60.err
python scripts/layer_prof.py --out gpurun_out/r2a/layer_prof.csv > gpurun_out/r2a/layer_prof.txt 2>&1; head -40 gpurun_out/r2a/layer_prof.txt
