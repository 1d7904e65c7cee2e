"""Debug: per-phase cycle accounting of tc_attn_kernel (CTA 0) at the nuScenes shape."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmtcoop_b200 import ops, _lib
lib = _lib.load()
dev = "cuda:0"
B, N_kv, L, H, Nq = 8, 56400, 1, 8, 900
q = (torch.randn(B, Nq, 256, device=dev) * 0.25).bfloat16()
k = torch.randn(B, L, H, N_kv, 32, device=dev).bfloat16()
vt = torch.randn(B, L, H, 32, N_kv, device=dev).bfloat16()
ops.cross_attn(q, k, vt, 0); torch.cuda.synchronize()
buf = torch.zeros(32 + 148, dtype=torch.int64, device=dev)
lib.cmt_debug_attn_timing.argtypes = [ctypes.c_void_p]
lib.cmt_debug_attn_timing(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.cross_attn(q, k, vt, 0); e1.record(); torch.cuda.synchronize()
lib.cmt_debug_attn_timing(ctypes.c_void_p(0))
t = buf.cpu().tolist()
tiles = max(t[6], 1)
print(f"kernel+merge {e0.elapsed_time(e1)*1e3:.0f} us; softmax tiles counted (2 warps): {tiles}")
names = {0: "softmax wait s_full", 1: "softmax ld S + arrive s_empty", 2: "softmax mask+max", 3: "softmax exps (incl. rare rescale)",
         4: "softmax wait pv_done", 5: "softmax st P + fence + arrive"}
for i, n in names.items():
    print(f"  {n:36s} {t[i]/tiles:8.0f} cyc/tile")
print(f"  softmax total                        {sum(t[0:6])/tiles:8.0f} cyc/tile")
it = tiles / 2
for i, n in {8: "mma wait k_full", 9: "mma wait s_empty0", 10: "mma wait s_empty1", 11: "mma wait v_full", 12: "mma wait p_full0",
             13: "mma wait p_full1", 15: "mma wait q_full(total)", 16: "tma wait k_empty", 17: "tma wait v_empty", 18: "tma wait q_empty(total)"}.items():
    print(f"  {n:36s} {t[i]/it:8.0f} cyc/iter")

import statistics
c = t[32:32 + 148]
print("per-CTA softmax-thread0 cycles: min %d median %d max %d; wall kernel %.0f us -> implied SM clock %.2f GHz (max-cycle CTA)" % (
    min(c), statistics.median(c), max(c), e0.elapsed_time(e1) * 1e3, max(c) / (e0.elapsed_time(e1) * 1e3) / 1e3))
print("first 8 CTAs:", c[:8])
