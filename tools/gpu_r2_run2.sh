#!/bin/bash
# round 2, run 2: fused shared_conv + segmented GEMM tests, whole suite, bench with the shared_conv scope
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder.py -m gpu -q --no-header -p no:cacheprovider -x -k "layernorm or zero_target" > gpurun_out/r2_tests_5a.log 2>&1
echo "new kernel tests exit=$? $(tail -n 2 gpurun_out/r2_tests_5a.log | tr '\n' ' ')"
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rA --durations=10 > gpurun_out/r2_tests_5.log 2>&1
echo "tests exit=$? $(tail -n 2 gpurun_out/r2_tests_5.log | tr '\n' ' ')"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_5.json 2> gpurun_out/r2_bench_5.err
echo "bench exit=$?"; tail -c 1800 gpurun_out/r2_bench_5.json; tail -5 gpurun_out/r2_bench_5.err
for wl in coop_lidar coop_fusion; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2_bench_5_$wl.json 2> gpurun_out/r2_bench_5_$wl.err
  echo "bench $wl exit=$?"; tail -3 gpurun_out/r2_bench_5_$wl.err
done
