#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -m gpu -q --no-header -p no:cacheprovider -x -k "attention or cross_attn or flash" > gpurun_out/attn_tests.log 2>&1
echo "attn tests exit=$? $(tail -n 2 gpurun_out/attn_tests.log | tr '\n' ' ')"
cp cmt-cooperative-perception_b200/libcmtcoop_b200.so build_variants/lib_new.so
bash tools/ab_bench.sh "$@"
