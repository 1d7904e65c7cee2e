#!/bin/bash
# round 2, run 1: whole GPU suite (incl. the BASELINE-shape parity tests) + bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rA --durations=15 > gpurun_out/r2_tests_1.log 2>&1
echo "tests exit=$? $(tail -n 2 gpurun_out/r2_tests_1.log | tr '\n' ' ')"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err
echo "bench exit=$?"; head -c 1500 gpurun_out/r2_bench_1.json
timeout 600 python bench.py --steps 20 --warmup 5 --feat-dtype fp32 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_1_fp32feat.json 2> gpurun_out/r2_bench_1_fp32feat.err
echo "bench fp32 feat exit=$?"
for wl in coop_lidar coop_fusion; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2_bench_1_$wl.json 2> gpurun_out/r2_bench_1_$wl.err
  echo "bench $wl exit=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 --workload lidar128 --batch 1 --no-cpu-baseline > gpurun_out/r2_bench_1_lidar128.json 2> gpurun_out/r2_bench_1_lidar128.err
echo "bench lidar128 exit=$?"
