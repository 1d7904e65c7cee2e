#!/bin/bash
# validation of the tree: peer-kernel tests (emulated group opted in), whole GPU suite, smoke, default bench
mkdir -p gpurun_out
CMT_TEST_PEER_EMULATION=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -rA -k "peer" > gpurun_out/r2_tests_peer_emulated.log 2>&1
echo "peer tests exit=$? $(tail -n 1 gpurun_out/r2_tests_peer_emulated.log)"
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rA --durations=8 > gpurun_out/r2_tests_final.log 2>&1
echo "tests exit=$? $(tail -n 2 gpurun_out/r2_tests_final.log | tr '\n' ' ')"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/r2_smoke.log)"
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit=$? $(python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1]); print('value %.1f ms %.3f e2e %.1f parity %s attn %.4f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['rel_l2'],d['roofline']['avg_launch_ms']))")"
