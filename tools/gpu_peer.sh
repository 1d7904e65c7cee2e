#!/bin/bash
# KV-token split: NCCL all-gather + merge against the peer-memory exchange + merge kernel, at N GPUs
# usage: gpu_peer.sh N [variants...]   variants: test peer peer0 peer1 nccl_graph nccl
N=${1:-2}; shift
mkdir -p gpurun_out
run() { # name, extra args
  name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-shared-conv-leg "$@" > gpurun_out/r2_${name}_$N.json 2> gpurun_out/r2_${name}_$N.err
  echo "$name N=$N exit=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/r2_${name}_$N.json').read().strip().splitlines()[-1]); print('value %.1f ms %.3f eager %.3f e2e %.1f parity %s'%(d['value'],d['ms_per_step'],d['config']['eager_ms_per_step'],d['e2e']['value'],d['parity'] and d['parity']['rel_l2']))
except Exception as e: print('ERR',e)
")"
  tail -2 gpurun_out/r2_${name}_$N.err | grep -v "OMP_NUM\|\*\*\*"
}
for v in "$@"; do
  case $v in
    test) timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider -rA > gpurun_out/r2_test_gpu_multi_peer.log 2>&1
          echo "test_gpu_multi exit=$? $(tail -n 2 gpurun_out/r2_test_gpu_multi_peer.log | tr '\n' ' ')";;
    peer) run kvsplit_peer --kv-split --kv-split-peer;;
    peer0) CMT_PEER_SCATTER=0 run kvsplit_peer_gather --kv-split --kv-split-peer;;
    peer1) CMT_PEER_SCATTER=1 run kvsplit_peer_scatter --kv-split --kv-split-peer;;
    nccl_graph) run kvsplit_graph --kv-split --kv-split-graph;;
    nccl) run kvsplit --kv-split;;
  esac
done
true
