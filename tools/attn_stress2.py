"""Back-to-back (no host sync, PDL on) attention launches at the decoder's two shapes; checks every result at the end."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmtcoop_b200 import ops
dev = "cuda:0"; H = 8
g = torch.Generator(device=dev).manual_seed(1)
def mk(B, Nq, N):
    q = (torch.randn(B, Nq, H * 32, generator=g, device=dev) * 0.2).bfloat16()
    k = torch.randn(B, 1, H, N, 32, generator=g, device=dev).bfloat16()
    ld = (N + 7) // 8 * 8
    vt = torch.zeros(B, 1, H, 32, ld, device=dev, dtype=torch.bfloat16); vt[..., :N] = torch.randn(B, 1, H, 32, N, generator=g, device=dev).bfloat16()
    qn = q.float().view(B, Nq, H, 32).pow(2).sum(-1).amax(1).contiguous(); kn = k.float().pow(2).sum(-1).amax(-1).contiguous()
    return q, k, vt, qn, kn
cases = [mk(8, 900, 56400), mk(8, 900, 900)]
refs = [ops.cross_attn(c[0], c[1], c[2], 0, o_dtype=torch.float32, simt=True) for c in cases]
torch.cuda.synchronize()
outs = []
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    for ci, (q, k, vt, qn, kn) in enumerate(cases):
        outs.append((ci, ops.cross_attn(q, k, vt, 0, o_dtype=torch.float32, q_norm2=qn, k_norm2=kn)))
        if ci == 1:
            outs.append((ci, ops.cross_attn(q, k, vt, 0, o_dtype=torch.float32)))   # self-attention runs without norms (online)
torch.cuda.synchronize()
bad = sum(1 for ci, o in outs if not (float((o - refs[ci]).norm() / refs[ci].norm()) < 5e-3))
print("launches", len(outs), "bad", bad)
sys.exit(1 if bad else 0)
