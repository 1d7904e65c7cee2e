#!/bin/bash
lib=cmt-cooperative-perception_b200/libcmtcoop_b200.so
cp $lib /tmp/lib_shipped.so
cp build_variants/lib_aa10.so $lib || exit 9
timeout 120 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -m gpu -q --no-header -p no:cacheprovider -x -k "attention or cross_attn or flash" 2>&1 | tail -1
for i in 1 2 3; do
  timeout 300 python tools/one_forward.py 600 > gpurun_out/fw_$i.log 2>&1; rc=$?; echo "aa10 run $i exit=$rc $(grep -c ok gpurun_out/fw_$i.log) x10 forwards ok"
done
cp /tmp/lib_shipped.so $lib
bash tools/ab_bench.sh shipped aa8 aa10 aa12 shipped aa10
