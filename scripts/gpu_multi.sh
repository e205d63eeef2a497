#!/bin/bash
# Multi-GPU pass: the tests that need two GPUs, then the C3 strong-scaling line on all visible GPUs.
N=$(python -c "import torch; print(torch.cuda.device_count())")
OUT=gpurun_out/r2m; mkdir -p $OUT
if [ "$N" -ge 2 ] && [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 600 python -m pytest -q -p no:cacheprovider tests/test_scene_gpu.py tests/test_train_gpu.py -k "two_gpu or two_rank or gpus" 2>&1 | tail -3
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload c3 --gpus $N --steps ${STEPS:-100} --warmup 3 --no-cpu --no-gpu-baseline > $OUT/bench_c3_n$N.json 2> $OUT/bench_c3_n$N.err
tail -c 400 $OUT/bench_c3_n$N.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_c3_n$N.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","n_gpus","steps","ms_per_step","non_sampling_ms","sampling_ms","tiles_this_rank","ideal_speedup","clocks")})
PY
