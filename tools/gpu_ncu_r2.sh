#!/bin/bash
# round 2 ncu captures (--set full, one launch each, after a plain run of the same command)
mkdir -p gpurun_out
python tools/prof_kernels.py attnstatic 3 > gpurun_out/ncu_plain_attn.log 2>&1; echo "plain attn exit=$? $(cat gpurun_out/ncu_plain_attn.log | tr '\n' ' ')"
python tools/prof_kernels.py conv 3 > gpurun_out/ncu_plain_conv.log 2>&1; echo "plain conv exit=$? $(cat gpurun_out/ncu_plain_conv.log | tr '\n' ' ')"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_attn_db_kernel --launch-skip 2 -c 1 -o gpurun_out/r2_attn_static -f python tools/prof_kernels.py attnstatic 3 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 1 -c 1 -o gpurun_out/r2_shared_conv -f python tools/prof_kernels.py conv 3 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv exit=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
