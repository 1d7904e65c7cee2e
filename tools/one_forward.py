"""N eager forwards of the bench workload (for compute-sanitizer / debugging)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cmtcoop_b200 import synth, ops
ops.FORCE_ONLINE_SOFTMAX = os.environ.get("CMT_FORCE_ONLINE") == "1"
from cmtcoop_b200.plugin import build_head
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kind, cfg, inputs = bench.build_case("nusc", B, seed=0)
head = build_head({k: v for k, v in cfg.items() if not k.startswith("_")})
synth.load_synth_weights(head, 0)
head = head.to("cuda:0").eval().set_precision("bf16")
head.apply_shared_conv = False
feats = {k: torch.from_numpy(v).bfloat16().to("cuda:0") for k, v in inputs.items() if isinstance(v, np.ndarray)}
with torch.no_grad():
    for i in range(n):
        rets = head.forward_single(feats["pts_feats"], feats["img_feats"], inputs["img_metas"])
        if i % 10 != 9:
            continue
        torch.cuda.synchronize()
        print("forward", i, "ok", float(rets[0]["cls_logits"].float().abs().mean()), flush=True)
