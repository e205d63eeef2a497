#!/bin/bash
# Round-2 evidence pass: sanitizers on the final code, ncu launch list of the bench command, ncu --set full of the dominant kernels.
OUT=gpurun_out/r2e
mkdir -p $OUT /tmp/ncu
SKIP='not (176 or 512 or T2000 or two_gpu or two_rank)'
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --log-file $OUT/memcheck.log python -m pytest -q -p no:cacheprovider tests -m gpu -k "$SKIP" > $OUT/memcheck.out 2>&1
echo "memcheck exit $?"; tail -3 $OUT/memcheck.out; grep -E "ERROR SUMMARY" $OUT/memcheck.log
echo "== racecheck (tensor-core conv / GEMM kernels, small UNet forward + sampling in bf16, training step)"
timeout 1500 compute-sanitizer --tool racecheck --log-file $OUT/racecheck.log python -m pytest -q -p no:cacheprovider tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_train_gpu.py -m gpu -k "(tc or dispatch or small or sampler or reference or loss) and $SKIP and not full128 and not wide64 and not full32 and not c4" > $OUT/racecheck.out 2>&1
echo "racecheck exit $?"; tail -3 $OUT/racecheck.out; grep -E "RACECHECK SUMMARY|ERROR SUMMARY" $OUT/racecheck.log
echo "== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-gpu-baseline > $OUT/bench_under_ncu.log 2>&1
echo "launches exit $?"; wc -l $OUT/bench_launches.csv
echo "== ncu --set full"
CMD="python scripts/step_time.py --precision bf16 --batches 176 --iters 1"
for spec in "h464:conv_halo_kernel<4, 64, 9, false>:4:3" "h2128:conv_halo_kernel<2, 128, 9, false>:2:2" "h1256p:conv_halo_kernel<1, 256, 9, true>:2:2" "fin:gn_finalize_kernel:4:2" "post:posterior_kernel:0:1"; do
  IFS=: read tag kre skip cnt <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$kre" -s $skip -c $cnt -o /tmp/ncu/$tag $CMD > $OUT/ncu_$tag.log 2>&1
  echo "$tag exit $?"
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > $OUT/$tag.raw.csv 2>/dev/null
done
# the posterior kernel only runs in the sampler: capture it from a short sampling run
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:posterior_kernel -s 3 -c 1 -o /tmp/ncu/post python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-gpu-baseline > $OUT/ncu_post.log 2>&1
ncu -i /tmp/ncu/post.ncu-rep --page raw --csv > $OUT/post.raw.csv 2>/dev/null
ls -la $OUT
