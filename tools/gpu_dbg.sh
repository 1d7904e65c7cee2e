#!/bin/bash
lib=cmt-cooperative-perception_b200/libcmtcoop_b200.so
cp $lib /tmp/lib_shipped.so
cp build_variants/lib_$1.so $lib
for i in $(seq 1 ${2:-8}); do
  timeout 200 python tools/bench_dbg.py > gpurun_out/bdbg_$i.log 2>&1; echo "run $i exit=$? $(grep 'BENCH\|records' gpurun_out/bdbg_$i.log | tr '\n' ' ' | cut -c1-200)"
  if grep -q "records: [1-9]" gpurun_out/bdbg_$i.log; then grep "blk" gpurun_out/bdbg_$i.log | head -70; break; fi
done
cp /tmp/lib_shipped.so $lib
