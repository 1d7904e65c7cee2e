#!/bin/bash
# plain bench run (the only numbers that count), then an ncu launch list of the SAME short command
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit=$?"; tail -n 3 gpurun_out/ncu.log
