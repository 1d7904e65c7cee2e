#!/bin/bash
# launch list (per-kernel durations, cold-cache/serialised) of one bench invocation: profiles/ evidence for the share of each kernel
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-parity --no-shared-conv-leg"
timeout 600 python bench.py $ARGS > gpurun_out/ll_plain.json 2> gpurun_out/ll_plain.err; echo "plain exit=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv python bench.py $ARGS > gpurun_out/ll_ncu.log 2>&1; echo "ncu exit=$?"
tail -3 gpurun_out/ll_ncu.log
