#!/bin/bash
lib=cmt-cooperative-perception_b200/libcmtcoop_b200.so
cp $lib /tmp/lib_shipped.so
for v in "$@"; do
  cp build_variants/lib_$v.so $lib
  for mode in "" "CMT_PDL=0"; do
    for i in 1 2 3; do
      env $mode timeout 200 python tools/attn_stress2.py 30 > gpurun_out/stress2_$v.log 2>&1; echo "$v [$mode] run $i exit=$? $(tail -1 gpurun_out/stress2_$v.log | cut -c1-150)"
    done
  done
done
cp /tmp/lib_shipped.so $lib
