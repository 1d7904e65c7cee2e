"""Stress of cmt_cross_attn_fwd (band-aligned / split-band scheduling): many shapes, both softmax instantiations, against the
CUDA-core comparator on the same bf16 operands.  usage: python tools/attn_stress.py [rounds]"""
import os, sys, math, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmtcoop_b200 import ops
dev = "cuda:0"
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shapes = [(8, 900, 56400), (1, 900, 5000), (2, 900, 16384), (1, 4, 64), (2, 130, 1000), (3, 257, 2049), (2, 388, 1500), (1, 1, 130),
          (1, 96, 300), (4, 900, 32400), (1, 384, 700), (2, 512, 900), (1, 640, 333), (1, 1000, 2000), (1, 33, 129), (8, 900, 7050)]
g = torch.Generator(device=dev).manual_seed(1)
bad = 0
for rd in range(rounds):
    for (B, Nq, N) in shapes:
        H = 8
        q = (torch.randn(B, Nq, H * 32, generator=g, device=dev) * 0.2).bfloat16()
        k = torch.randn(B, 1, H, N, 32, generator=g, device=dev).bfloat16()
        ld = (N + 7) // 8 * 8
        vt = torch.zeros(B, 1, H, 32, ld, device=dev, dtype=torch.bfloat16)
        vt[..., :N] = torch.randn(B, 1, H, 32, N, generator=g, device=dev).bfloat16()
        qn = q.float().view(B, Nq, H, 32).pow(2).sum(-1).amax(1).contiguous()
        kn = k.float().pow(2).sum(-1).amax(-1).contiguous()
        ref = ops.cross_attn(q, k, vt, 0, o_dtype=torch.float32, simt=True)
        for static in (True, False):
            t0 = time.perf_counter()
            o = ops.cross_attn(q, k, vt, 0, o_dtype=torch.float32, q_norm2=qn if static else None, k_norm2=kn if static else None)
            torch.cuda.synchronize()
            rel = float((o - ref).norm() / ref.norm())
            ok = rel < 5e-3 and bool(torch.isfinite(o).all())
            bad += 0 if ok else 1
            print(f"round {rd} B={B} Nq={Nq} N={N} static={static}: rel {rel:.2e} {'ok' if ok else 'BAD'} ({(time.perf_counter()-t0)*1e3:.1f} ms)", flush=True)
print("bad:", bad)
sys.exit(1 if bad else 0)
