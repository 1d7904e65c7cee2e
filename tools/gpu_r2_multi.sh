#!/bin/bash
# multi-GPU run: 2-GPU parity test of the KV-token split, then bench lines: frame sharding and KV split at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider -rA > gpurun_out/r2_test_gpu_multi.log 2>&1
  echo "test_gpu_multi exit=$? $(tail -n 2 gpurun_out/r2_test_gpu_multi.log | tr '\n' ' ')"
fi
run() { # name, extra args
  name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-shared-conv-leg "$@" > gpurun_out/r2_${name}_$N.json 2> gpurun_out/r2_${name}_$N.err
  echo "$name N=$N exit=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/r2_${name}_$N.json').read().strip().splitlines()[-1]); print('value %.1f ms %.3f e2e %.1f parity %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity'] and d['parity']['rel_l2']))
except Exception as e: print('ERR',e)
")"
  tail -2 gpurun_out/r2_${name}_$N.err
}
run kvsplit --kv-split
[ "$N" = "2" ] && run kvsplit_graph --kv-split --kv-split-graph
run shard
# host->device ceiling of this box: every rank copies 133 MB of pinned memory at the same time (no compute)
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/e2e_probe2.py 2>&1 | grep "rank" | sort > gpurun_out/r2_h2d_probe_$N.txt
echo "h2d probe N=$N: $(grep -c 'H2D 133MB' gpurun_out/r2_h2d_probe_$N.txt) lines; $(grep 'H2D 133MB' gpurun_out/r2_h2d_probe_$N.txt | tail -1)"
