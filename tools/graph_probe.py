"""Probe: CUDA-graph replay of CmtHead.forward_single vs eager launches at the bench shape (does the step have launch gaps?)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cmtcoop_b200 import synth
from cmtcoop_b200.plugin import build_head
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
kind, cfg, inputs = bench.build_case("nusc", B)
head = build_head({k: v for k, v in cfg.items() if not k.startswith("_")})
synth.load_synth_weights(head, 0)
head = head.to(dev).eval().set_precision("bf16")
head.apply_shared_conv = False
res = {k: torch.from_numpy(v).to(dev) for k, v in inputs.items() if isinstance(v, np.ndarray)}
metas = inputs["img_metas"]
fwd = lambda: head.forward_single(res["pts_feats"], res["img_feats"], metas)
def timed(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    for _ in range(5): fwd()
    print(f"eager : {timed(fwd):.3f} ms/step")
    from cmtcoop_b200.runtime import GraphedForward
    g = GraphedForward(head, metas, res)
    out = g()
    ref = fwd()
    worst = max(float((out[0][n] - ref[0][n]).abs().max()) for n in ref[0])
    print(f"graph vs eager max abs diff {worst:.3e}")
    print(f"graph : {timed(lambda: g()):.3f} ms/step")
