#!/bin/bash
lib=cmt-cooperative-perception_b200/libcmtcoop_b200.so
cp $lib /tmp/lib_shipped.so
cp build_variants/lib_$1.so $lib
timeout 400 python tools/dbg_loop.py 400 > gpurun_out/dbg_loop.log 2>&1; echo "exit=$?"; tail -70 gpurun_out/dbg_loop.log | cut -c1-200
cp /tmp/lib_shipped.so $lib
