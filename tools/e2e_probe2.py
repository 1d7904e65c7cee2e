"""torchrun probe: where does the e2e loop spend its time at N ranks? (H2D bandwidth per rank, host time per phase)"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
use_nccl = "--no-nccl" not in sys.argv
if world > 1:
    dist.init_process_group("nccl" if use_nccl else "gloo", **({"device_id": dev} if use_nccl else {}))
def log(*a):
    print(f"[rank {rank}]", *a, flush=True)
log("OMP", os.environ.get("OMP_NUM_THREADS"), "affinity", len(os.sched_getaffinity(0)), "cpus")
x = torch.randn(8, 256, 180, 180).bfloat16()
xp = x.pin_memory()
d = torch.empty_like(xp, device=dev)
s_in = torch.cuda.Stream(dev)
for it in range(4):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(s_in):
        e0.record(); d.copy_(xp, non_blocking=True); e1.record()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    log(f"H2D 133MB pinned: enqueue {1e3*(t1-t0):.2f} ms, wall {1e3*(t2-t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms -> {x.numel()*2/e0.elapsed_time(e1)/1e6:.1f} GB/s")
# compute + copy overlap: a long kernel on the current stream while copying
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for it in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_in):
        e0.record(); d.copy_(xp, non_blocking=True); e1.record()
    for _ in range(10): a @ a
    torch.cuda.synchronize()
    log(f"H2D under matmuls: device {e0.elapsed_time(e1):.2f} ms")
if world > 1: dist.destroy_process_group()
