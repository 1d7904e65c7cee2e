"""What separates the end-to-end step (PipelinedRunner) from the resident one at one GPU?  Times 20 steps of
  resident      : graph replays back to back
  interference  : the same replays while another stream copies the step's 231 MB of pinned features in a loop (no dependency)
  no_copy_in    : PipelinedRunner with the host->device copy removed (events, graph per slot and device->host copies stay)
  runner        : PipelinedRunner as shipped."""
import os
import sys
import threading

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cmtcoop_b200 import synth  # noqa: E402
from cmtcoop_b200.plugin import build_head  # noqa: E402
from cmtcoop_b200.runtime import GraphedForward, PipelinedRunner  # noqa: E402

dev = torch.device("cuda:0")
B, steps = 8, 20
kind, cfg, inputs = bench.build_case("nusc", B, seed=0)
head = build_head({k: v for k, v in cfg.items() if not k.startswith("_")})
synth.load_synth_weights(head, 0)
head = head.to(dev).eval().set_precision("bf16")
head.apply_shared_conv = False
keys = [k for k, v in inputs.items() if isinstance(v, np.ndarray)]
host = {k: torch.from_numpy(inputs[k]).to(torch.bfloat16).pin_memory() for k in keys}
resident = {k: v.to(dev) for k, v in host.items()}
metas = inputs["img_metas"]


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


with torch.no_grad():
    for _ in range(3):
        head.forward_single(resident["pts_feats"], resident["img_feats"], metas)
    g = GraphedForward(head, metas, resident, adopt_inputs=True)
    for _ in range(3):
        g()
    print("resident      %.3f ms/step" % timed(lambda: [g() for _ in range(steps)]))

    side = torch.cuda.Stream()
    scratch = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    stop = False

    def pump():
        with torch.cuda.stream(side):
            while not stop:
                for k in keys:
                    scratch[k].copy_(host[k], non_blocking=True)
                side.synchronize()
    th = threading.Thread(target=pump)
    th.start()
    print("interference  %.3f ms/step" % timed(lambda: [g() for _ in range(steps)]))
    print("interference  %.3f ms/step" % timed(lambda: [g() for _ in range(steps)]))
    stop = True
    th.join()

    runner = PipelinedRunner(head, metas, host, dev, use_cuda_graph=True)
    runner.run([host] * 3)
    print("runner        %.3f ms/step" % timed(lambda: runner.run([host] * steps)))
    print("runner        %.3f ms/step" % timed(lambda: runner.run([host] * steps)))
    orig = runner._enqueue_copy_in

    def no_copy(slot, host_inputs):
        with torch.cuda.stream(runner.s_in):
            if runner._primed[slot]:
                runner.s_in.wait_event(runner.ev_free[slot])
            runner.ev_in[slot].record(runner.s_in)
    runner._enqueue_copy_in = no_copy
    print("no_copy_in    %.3f ms/step" % timed(lambda: runner.run([host] * steps)))
    runner._enqueue_copy_in = orig
    runner.keep_results = False
    print("runner, results left in the staging buffers %.3f ms/step" % timed(lambda: runner.run([host] * steps)))
