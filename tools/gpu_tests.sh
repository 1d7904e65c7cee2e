#!/bin/bash
# Runs the GPU parity suites one file/group at a time under a timeout so that a fault in one
# kernel family cannot take the others (or the box) down with it.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() {
  name=$1; shift
  echo "=== $name"
  timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider "$@" > gpurun_out/test_$name.log 2>&1
  echo "exit=$? $(tail -n 3 gpurun_out/test_$name.log | tr '\n' ' ')"
}
run simple tests/test_gpu_kernels.py -k "not gemm and not attention"
run gemm tests/test_gpu_kernels.py -k "gemm"
run attn tests/test_gpu_kernels.py -k "attention"
run heads tests/test_gpu_heads.py
run decoder tests/test_gpu_decoder.py
run multi tests/test_gpu_multi.py
