"""Debug build (CMT_TRAP_REPORT): run forwards until a bounded wait times out, then print the host-mapped records."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cmtcoop_b200 import synth, _lib
from cmtcoop_b200.plugin import build_head
lib = _lib.load()
rec = torch.zeros(1 + 4 * 500, dtype=torch.int64).pin_memory()
lib.cmt_debug_attn_timing(ctypes.c_void_p(rec.data_ptr()))
kind, cfg, inputs = bench.build_case("nusc", 8, seed=0)
head = build_head({k: v for k, v in cfg.items() if not k.startswith("_")})
synth.load_synth_weights(head, 0)
head = head.to("cuda:0").eval().set_precision("bf16")
head.apply_shared_conv = False
feats = {k: torch.from_numpy(v).bfloat16().to("cuda:0") for k, v in inputs.items() if isinstance(v, np.ndarray)}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
try:
    with torch.no_grad():
        for i in range(n):
            head.forward_single(feats["pts_feats"], feats["img_feats"], inputs["img_metas"])
            if i % 10 == 9:
                torch.cuda.synchronize()
                print("forwards", i + 1, flush=True)
    torch.cuda.synchronize()
    print("no failure in", n, "forwards")
except Exception as e:
    print("FAILED:", str(e)[:100])
k = int(rec[0])
print("records:", k)
for j in range(min(k, 60)):
    a, off, par, line = (int(rec[1 + 4 * j + t]) for t in range(4))
    print(f"  blk {a & 0xffff} warp {a >> 16} bar_off 0x{off:x} parity {par} line {line}")
