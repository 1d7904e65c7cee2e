"""Host-side timing of the e2e loop pieces (why is PipelinedRunner slow under torchrun's OMP_NUM_THREADS=1?)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print("OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"), "torch threads", torch.get_num_threads())
dev = torch.device("cuda:0")
x = torch.randn(8, 256, 180, 180).bfloat16()
t0 = time.perf_counter(); xp = x.pin_memory(); t1 = time.perf_counter()
print("pin_memory 133MB: %.1f ms, is_pinned %s" % ((t1 - t0) * 1e3, xp.is_pinned()))
d = torch.empty_like(xp, device=dev)
for name, src in (("pinned", xp), ("pageable", x)):
    torch.cuda.synchronize()
    for _ in range(2):
        t0 = time.perf_counter(); d.copy_(src, non_blocking=True); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"H2D {name}: enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms -> {x.numel()*2/(t2-t0)/1e9:.1f} GB/s")
o = torch.randn(6, 8, 900, 10, device=dev)
h = torch.empty(o.shape).pin_memory()
t0 = time.perf_counter(); h.copy_(o, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
c = h.clone(); t2 = time.perf_counter()
print("D2H 1.7MB %.2f ms; clone %.2f ms pinned_clone=%s" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), c.is_pinned()))
big = torch.empty(6, 8, 900, 20).pin_memory()
t0 = time.perf_counter(); cs = [big.clone() for _ in range(6)]; t1 = time.perf_counter()
print("6 clones of 3.4MB pinned: %.2f ms" % (1e3 * (t1 - t0)))
