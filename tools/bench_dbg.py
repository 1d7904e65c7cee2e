"""Runs bench.main() under a CMT_TRAP_REPORT build with a HOST-mapped record buffer for bounded-wait timeouts."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtcoop_b200 import _lib
lib = _lib.load()
rec = torch.zeros(1 + 4 * 500, dtype=torch.int64).pin_memory()
real = lib.cmt_debug_attn_timing
real(ctypes.c_void_p(rec.data_ptr()))
class _Noop:
    argtypes = None
    def __call__(self, *a):
        return 0
lib.cmt_debug_attn_timing = _Noop()      # bench's own diagnostic calls must not replace the record buffer
import bench
sys.argv = ["bench.py", "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--no-parity", "--no-shared-conv-leg"]
try:
    bench.main()
    print("BENCH OK")
except BaseException as e:
    print("BENCH FAILED:", type(e).__name__, str(e)[:120])
k = int(rec[0])
print("records:", k)
for j in range(min(k, 80)):
    a, off, par, line = (int(rec[1 + 4 * j + t]) for t in range(4))
    print(f"  blk {a & 0xffff} warp {a >> 16} bar_off 0x{off:x} parity {par} line {line}")
os._exit(0)
