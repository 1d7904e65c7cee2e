"""Debug: clock64 event trace of CTA 0 of the three-warpgroup attention kernel at the nuScenes shape.
Per warpgroup and tile-step: S ready -> S in registers -> max done -> P stored; issuer: P seen -> next S issued."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmtcoop_b200 import ops, _lib
lib = _lib.load()
dev = "cuda:0"
B, N_kv, L, H, Nq = 8, 56400, 1, 8, 900
STEPS = 96
q = (torch.randn(B, Nq, 256, device=dev) * 0.25).bfloat16()
k = torch.randn(B, L, H, N_kv, 32, device=dev).bfloat16()
vt = torch.randn(B, L, H, 32, N_kv, device=dev).bfloat16()
ops.cross_attn(q, k, vt, 0); torch.cuda.synchronize()
buf = torch.zeros(3 * STEPS * 8 + 148, dtype=torch.int64, device=dev)
lib.cmt_debug_attn_timing.argtypes = [ctypes.c_void_p]
lib.cmt_debug_attn_timing(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.cross_attn(q, k, vt, 0); e1.record(); torch.cuda.synchronize()
lib.cmt_debug_attn_timing(ctypes.c_void_p(0))
cyc = buf[3 * STEPS * 8:].cpu().tolist()
us = e0.elapsed_time(e1) * 1e3
print(f'kernel+merge {us:.0f} us; per-CTA cycles min {min(cyc)} median {sorted(cyc)[74]} max {max(cyc)} -> SM clock >= {max(cyc) / us / 1e3:.2f} GHz')
t = buf[:3 * STEPS * 8].cpu().view(3, STEPS, 8)
t0 = int(t[0, 0, 0])
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 56
print("step wg |  S_ready   ld_done  max_done  P_stored | iss_P_seen iss_S_issued | wait_S  ld   max   exp  | step_period")
for st in range(lo, hi):
    for wg in range(3):
        e = [int(x) - t0 for x in t[wg, st]]
        prev = int(t[wg, st - 1, 0]) - t0 if st > 0 else 0
        prevP = int(t[wg, st - 1, 3]) - t0 if st > 0 else 0
        print(f"{st:4d} {wg:2d} | {e[0]:8d} {e[1]:8d} {e[2]:8d} {e[3]:8d} | {e[4]:9d} {e[5]:9d} | "
              f"{e[0]-prevP:6d} {e[1]-e[0]:4d} {e[2]-e[1]:5d} {e[3]-e[2]:5d} | {e[0]-prev:6d} | issuer: loop_top {e[6]:8d} wait {e[7]-e[6]:4d} fence {e[4]-e[7]:4d} issue {e[5]-e[4]:4d}")
import statistics
for wg in range(3):
    per = [int(t[wg, s, 0] - t[wg, s - 1, 0]) for s in range(20, STEPS)]
    ex = [int(t[wg, s, 3] - t[wg, s, 2]) for s in range(20, STEPS)]
    n_ = [int(t[wg, s, 0] - t[wg, s - 1, 3]) for s in range(20, STEPS)]
    iss = [int(t[wg, s, 4] - t[wg, s, 3]) for s in range(20, STEPS)]
    print(f"wg{wg}: period median {statistics.median(per)}, exps {statistics.median(ex)}, P_stored->next S_ready {statistics.median(n_)}, "
          f"P_stored->issuer saw it {statistics.median(iss)}")
