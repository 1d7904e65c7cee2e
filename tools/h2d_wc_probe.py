"""torchrun probe: host->device rate per rank when every rank copies at once, for three kinds of pinned host memory:
torch's pin_memory() (cudaHostAlloc default), cudaHostAlloc(portable), cudaHostAlloc(write-combined)."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.zeros(1, device=dev)
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
cudart.cudaHostAlloc.restype = ctypes.c_int


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), nbytes, flags)
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


nbytes = 231 * 1000 * 1000
src = torch.randint(0, 255, (nbytes,), dtype=torch.uint8)
kinds = {"torch pin_memory": lambda: src.pin_memory(),
         "cudaHostAlloc portable": lambda: host_alloc(nbytes, 1).copy_(src),
         "cudaHostAlloc write-combined": lambda: host_alloc(nbytes, 4).copy_(src)}
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s_in = torch.cuda.Stream(dev)
for name, make in kinds.items():
    h = make()
    pinned = h.is_pinned()
    rates = []
    for it in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s_in):
            e0.record()
            for _ in range(10):
                d.copy_(h, non_blocking=True)
            e1.record()
        torch.cuda.synchronize()
        rates.append(10 * nbytes / e0.elapsed_time(e1) / 1e6)
    ok = bool((d[:1000].cpu() == src[:1000]).all())
    print(f"[rank {rank}/{world}] {name}: is_pinned {pinned}, 10 x 231 MB: " + " / ".join(f"{r:.1f}" for r in rates) + f" GB/s, data ok {ok}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
