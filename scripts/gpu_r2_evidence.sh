#!/bin/bash
# Round-2 evidence pass on the final code: smoke, bench lines, sanitizers, ncu launch list of the bench command,
# ncu --set full of the dominant kernels, latency protocol, layer profiles.
OUT=gpurun_out/${1:-r2z}
mkdir -p $OUT /tmp/ncu
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== GPU test suite, one process, as the driver runs it"
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=8 > $OUT/tests.log 2>&1; echo "exit $?"; tail -14 $OUT/tests.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== bench (default: C2, full T = 2000 pass, e2e, cpu and gpu-library baselines)"
timeout 900 python bench.py > $OUT/bench_full.json 2> $OUT/bench_full.err; echo "exit $?"
echo "== bench c1 / c4"
timeout 300 python bench.py --workload c1 > $OUT/bench_c1.json 2> $OUT/bench_c1.err; echo "exit $?"
timeout 300 python bench.py --workload c4 --steps 20 --no-cpu --no-gpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "exit $?"
echo "== latency protocol"; STEP_LAT_N=1,5,11,44,176 timeout 300 python scripts/step_latency.py > $OUT/step_latency.jsonl 2>&1; cat $OUT/step_latency.jsonl
echo "== layer profiles"; timeout 300 python scripts/layer_prof.py --out $OUT/layer_prof.csv > $OUT/layer_prof.txt 2>&1; head -3 $OUT/layer_prof.txt
timeout 300 python scripts/layer_prof.py --n 5 --out $OUT/layer_prof_n5.csv > $OUT/layer_prof_n5.txt 2>&1; head -2 $OUT/layer_prof_n5.txt
# memcheck over the kernel, network, pre/post and scene tests (the long-chain, 176-latent, 512x512 and two-GPU cases left out)
SKIP='not (176 or 512 or T2000 or two_gpu or two_rank or train or dropin or e2e)'
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --log-file $OUT/memcheck.log python -m pytest -q -p no:cacheprovider tests -m gpu -k "$SKIP" > $OUT/memcheck.out 2>&1
echo "memcheck exit $?"; tail -2 $OUT/memcheck.out; grep -E "ERROR SUMMARY" $OUT/memcheck.log
echo "== racecheck (one bf16 UNet forward through the tensor-core kernels, and the tensor-core conv tests)"
timeout 540 compute-sanitizer --tool racecheck --log-file $OUT/racecheck.log python -m pytest -q -p no:cacheprovider tests/test_unet_gpu.py -k "eps_matches_reference and full32 and bf16" > $OUT/racecheck.out 2>&1
echo "racecheck exit $?"; tail -2 $OUT/racecheck.out; grep -E "RACECHECK SUMMARY|hazard" $OUT/racecheck.log | sort | uniq -c | head -5
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-gpu-baseline > $OUT/bench_under_ncu.log 2>&1
echo "launches exit $?"; wc -l $OUT/bench_launches.csv
echo "== ncu --set full"
CMD="python scripts/step_time.py --precision bf16 --batches 176 --iters 1"
for spec in "h464|conv_halo_kernel<.int.4, .int.64, .int.9, .bool.0>|4|3" "h2128|conv_halo_kernel<.int.2, .int.128, .int.9, .bool.0>|2|2" "h1256p|conv_halo_kernel<.int.1, .int.256, .int.9, .bool.1>|2|2" "pertap|conv_tc_kernel<.int.256>|2|2" "applyt|gn_apply_t_kernel|1|1" "flash|attn_flash_kernel|1|1" "i2c|im2col_small_kernel|0|1" "fin|conv_halo_kernel<.int.4, .int.16, .int.9, .bool.0>|0|1"; do
  IFS='|' read tag kre skip cnt <<< "$spec"
  timeout 420 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$kre" -s $skip -c $cnt -o /tmp/ncu/$tag $CMD > $OUT/ncu_$tag.log 2>&1
  echo "$tag exit $?"
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > $OUT/$tag.raw.csv 2>/dev/null
done
cp gpurun_out/r2f/post.raw.csv $OUT/post.raw.csv 2>/dev/null   # posterior_kernel is unchanged since that capture
ls -la $OUT
