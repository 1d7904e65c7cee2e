"""Launch each hot kernel a few times at the nuScenes-multimodal shape (B=8) for an ncu capture.
usage: python tools/prof_kernels.py [kproj|vproj|mlp1|mlp2|attn|attnonly|attnstatic|gather|raype|all]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmtcoop_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda:0"
B, N_kv, C, L, H, Nq = 8, 56400, 256, 6, 8, 900
torch.manual_seed(0)


def timed(name, fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per launch", flush=True)


if which in ("kproj", "vproj", "attn", "all"):
    x = torch.randn(B, N_kv, C, device=dev).bfloat16()
    w = (torch.randn(L * H * 32, C, device=dev) / 16).bfloat16()
    b = torch.randn(L * H * 32, device=dev)
    k = torch.empty((B, L, H, N_kv, 32), dtype=torch.bfloat16, device=dev)
    vt = torch.empty((B, L, H, 32, N_kv), dtype=torch.bfloat16, device=dev)
    if which in ("kproj", "all", "attn"):
        timed("kproj", lambda: ops.project_keys(x, w, b, L, H, out=k))
    if which in ("vproj", "all", "attn"):
        timed("vproj", lambda: ops.project_values_t(x, w, b, L, H, out=vt))
    if which in ("attn", "all"):
        q = (torch.randn(B, Nq, 256, device=dev) * 0.25).bfloat16()
        timed("attn", lambda: ops.cross_attn(q, k, vt, 2))
if which == "attnonly":   # attention alone on random K / V^T (A/B runs of kernel variants)
    k = torch.randn(B, 1, H, N_kv, 32, device=dev).bfloat16()
    vt = torch.randn(B, 1, H, 32, N_kv, device=dev).bfloat16()
    q = (torch.randn(B, Nq, 256, device=dev) * 0.25).bfloat16()
    timed("attn", lambda: ops.cross_attn(q, k, vt, 0))
    timed("attn", lambda: ops.cross_attn(q, k, vt, 0))
    # static softmax shift: score bound from the operand norms (what the projection epilogues provide)
    qn2 = q.float().view(B, Nq, H, 32).pow(2).sum(-1).amax(1).contiguous()
    kn2 = k.float().pow(2).sum(-1).amax(-1).contiguous()
    print("score bound (log2 units): max %.1f" % float((qn2 * kn2[:, 0]).sqrt().max()), flush=True)
    timed("attn_static", lambda: ops.cross_attn(q, k, vt, 0, q_norm2=qn2, k_norm2=kn2))
    # per-CTA cycle counts of one launch -> SM clock under this kernel (power-limited, well below the 1965 MHz maximum)
    import ctypes
    from cmtcoop_b200 import _lib
    lib = _lib.load()
    lib.cmt_debug_attn_timing.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(3 * 96 * 16 + 148, dtype=torch.int64, device=dev)
    lib.cmt_debug_attn_timing(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.cross_attn(q, k, vt, 0); e1.record(); torch.cuda.synchronize()
    lib.cmt_debug_attn_timing(ctypes.c_void_p(0))
    cyc = buf[3 * 96 * 16:].cpu().tolist()
    us = e0.elapsed_time(e1) * 1e3
    print(f"cycles/CTA median {sorted(cyc)[74]} max {max(cyc)}; {us:.0f} us -> SM clock >= {max(cyc) / us / 1e3:.3f} GHz", flush=True)
if which == "attnstatic":   # only static-shift attention launches (ncu capture of the shipped bench kernel)
    k = torch.randn(B, 1, H, N_kv, 32, device=dev).bfloat16()
    vt = torch.randn(B, 1, H, 32, N_kv, device=dev).bfloat16()
    q = (torch.randn(B, Nq, 256, device=dev) * 0.25).bfloat16()
    qn2 = q.float().view(B, Nq, H, 32).pow(2).sum(-1).amax(1).contiguous()
    kn2 = k.float().pow(2).sum(-1).amax(-1).contiguous()
    timed("attn_static", lambda: ops.cross_attn(q, k, vt, 0, q_norm2=qn2, k_norm2=kn2))
if which in ("mlp1", "mlp2", "all"):
    M = 8 * 6 * 4000
    a = torch.randn(M, 192, device=dev).bfloat16()
    w0 = (torch.randn(1024, 192, device=dev) / 14).bfloat16()
    b0 = torch.randn(1024, device=dev)
    w1 = (torch.randn(256, 1024, device=dev) / 32).bfloat16()
    b1 = torch.randn(256, device=dev)
    h = ops.linear(a, w0, b0, relu=True)
    if which in ("mlp1", "all"):
        timed("mlp1", lambda: ops.linear(a, w0, b0, relu=True))
    if which in ("mlp2", "all"):
        timed("mlp2", lambda: ops.linear(h, w1, b1, out_dtype=torch.float32))
if which in ("gather", "all"):
    xb = torch.randn(B, 256, 180, 180, device=dev)
    xi = torch.randn(B * 6, 256, 40, 100, device=dev)
    bp = torch.randn(32400, 256, device=dev)
    rp = torch.randn(B * 6 * 4000, 256, device=dev)
    timed("gather", lambda: ops.gather_tokens(xb, xi, bp, rp, B, 6))
if which in ("raype", "all"):
    m = torch.eye(4, device=dev).repeat(B * 6, 1, 1).contiguous()
    timed("raype", lambda: ops.ray_pe(m, 40, 100, 64, 640.0, 1600.0, [-54, -54, -5, 54, 54, 3]))
    # 64 frames (590 MB of bf16 output: beyond the 126 MB L2, so the time is an HBM write time)
    m64 = torch.eye(4, device=dev).repeat(64 * 6, 1, 1).contiguous()
    timed("raype_b64", lambda: ops.ray_pe(m64, 40, 100, 64, 640.0, 1600.0, [-54, -54, -5, 54, 54, 3]))
if which == "conv":   # shared_conv implicit GEMM at the config-3 shape (8 frames, 512 -> 256 channels, 180 x 180)
    x = torch.randn(B, 512, 180, 180, device=dev).bfloat16()
    w = (torch.randn(256, 9 * 512, device=dev) / 68).bfloat16()
    bias = torch.randn(256, device=dev)
    pos = torch.randn(180 * 180, 256, device=dev)
    xk = torch.empty((B, N_kv, 256), dtype=torch.bfloat16, device=dev)
    xv = torch.empty_like(xk)
    xp = ops.nchw_to_padded_nhwc(x)
    timed("nchw_to_padded_nhwc", lambda: ops.nchw_to_padded_nhwc(x, xp))
    timed("shared_conv", lambda: ops.shared_conv_tokens(xp, w, bias, pos, xk, xv, 180, 180))
