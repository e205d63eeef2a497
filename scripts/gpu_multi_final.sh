#!/bin/bash
# Final 2-GPU pass: the three tests that need two GPUs, and the bench through torchrun as the driver launches it.
OUT=gpurun_out/r2z3; mkdir -p $OUT
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_scene_gpu.py tests/test_train_gpu.py -k "two_gpu or two_rank" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-gpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "exit $?"
tail -1 $OUT/bench_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['value'], d['scaling'], d.get('warmup_note'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-160
