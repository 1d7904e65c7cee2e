import sys, math, torch
sys.path.insert(0, '/root/repo')
from cmtcoop_b200 import ops
sys.path.insert(0, '/root/repo/tests')
from test_gpu_kernels import _attn_ref
for (B, Nq, N_kv) in [(3, 257, 2049), (1, 900, 5000), (2, 130, 1000)]:
    H = 8
    g = torch.Generator().manual_seed(Nq + N_kv)
    q = (torch.randn(B, Nq, H * 32, generator=g) * (ops.LOG2E / math.sqrt(32)) * 2.0).bfloat16()
    k = torch.randn(B, 1, H, N_kv, 32, generator=g).bfloat16()
    ld = (N_kv + 7) // 8 * 8
    vt = torch.zeros(B, 1, H, 32, ld); vt[..., :N_kv] = torch.randn(B, 1, H, 32, N_kv, generator=g); vt = vt.bfloat16()
    wo, wl = _attn_ref(q, k, vt, N_kv, 0, N_kv)
    o, lse = ops.cross_attn(q.cuda(), k.cuda(), vt.cuda(), 0, o_dtype=torch.float32, return_lse=True)
    d = (lse.cpu() - wl).abs()
    print(B, Nq, N_kv, "o rel", float((o.cpu().double() - wo).norm() / wo.norm()), "lse max", float(d.max()), "mean", float(d.mean()),
          "argmax", [int(x) for x in torch.unravel_index(d.argmax(), d.shape)])
