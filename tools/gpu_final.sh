#!/bin/bash
# final validation of the committed tree: whole GPU suite, smoke, default bench, the other workloads, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rA --durations=8 > gpurun_out/r2_tests_final.log 2>&1
echo "tests exit=$? $(tail -n 2 gpurun_out/r2_tests_final.log | tr '\n' ' ')"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/r2_smoke.log)"
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit=$?"
for wl in coop_lidar coop_fusion; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2_bench_final_$wl.json 2> gpurun_out/r2_bench_final_$wl.err; echo "bench $wl exit=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 --workload lidar128 --batch 1 --no-cpu-baseline > gpurun_out/r2_bench_final_lidar128.json 2> gpurun_out/r2_bench_final_lidar128.err; echo "bench lidar128 exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --batch 64 --no-cpu-baseline --no-shared-conv-leg > gpurun_out/r2_bench_final_b64.json 2> gpurun_out/r2_bench_final_b64.err; echo "bench b64 exit=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference exit=$?"
