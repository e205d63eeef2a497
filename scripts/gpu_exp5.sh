#!/bin/bash
mkdir -p gpurun_out/exp5
timeout 900 python -m pytest -q --tb=short -p no:cacheprovider tests/test_train_gpu.py tests/test_scene_gpu.py -s 2>&1 | tail -12
python bench.py --workload c5 --steps 5 > gpurun_out/exp5/bench_c5_n1.json 2> gpurun_out/exp5/bench_c5_n1.err; echo "c5 n1 exit $?"; tail -c 400 gpurun_out/exp5/bench_c5_n1.err; cat gpurun_out/exp5/bench_c5_n1.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c5 --steps 5 > gpurun_out/exp5/bench_c5_n2.json 2> gpurun_out/exp5/bench_c5_n2.err; echo "c5 n2 exit $?"; tail -c 400 gpurun_out/exp5/bench_c5_n2.err; cat gpurun_out/exp5/bench_c5_n2.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c3 --steps 40 --no-cpu > gpurun_out/exp5/bench_c3_n2.json 2> gpurun_out/exp5/bench_c3_n2.err; echo "c3 n2 exit $?"; tail -c 400 gpurun_out/exp5/bench_c3_n2.err; cat gpurun_out/exp5/bench_c3_n2.json | cut -c1-300
