#!/usr/bin/env python
"""bench.py -- CmtTransformer+PE forward throughput (frames/s), the metric BASELINE.json names.

    python bench.py --gpus N --steps K --warmup W [--workload nusc|coop_lidar|coop_fusion|lidar128] [--batch B]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port), same metric/config

A "step" is one forward of the hot path over one batch of synthetic frames: camera-ray PE, BEV PE,
query embeddings, token gather, all-layer K/V projection, the 6-layer decoder (cross-attention on the
tcgen05 flash kernel) and the task heads -- i.e. CmtHead.forward_single from the post-shared_conv BEV
map and the neck's image features (SURVEY.md 8(d) scope).  Default workload: BASELINE.json configs[2]/[4]
shape, nuScenes multimodal (6 cams 40x100 tokens + 180x180 BEV = 56 400 K/V tokens, 900 queries, 6 layers,
d=256, bf16), 8 frames per GPU; frames shard across GPUs with no data-path collective (weak scaling).

One JSON line on stdout (rank 0).  `value` = device-timed throughput with inputs resident in HBM;
`e2e` = the same forward through the public head API from pinned HOST buffers (H2D + D2H in the timed
region); `roofline` is for the dominant kernel (tc_attn_db_kernel); `cpu_baseline` is the CPU oracle on a
bounded sample.  Only the cpu_baseline / --impl reference legs touch oracle/.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cmtcoop_b200 import synth  # noqa: E402

METRIC = "CmtTransformer+PE fwd frames/sec"
WORKLOADS = {
    # name: (head kind, bev_hw, n_views, description)
    "nusc": ("CmtHead", 180, 6, "CMT nuScenes multimodal: 6 cams 1600x640 stride-16 (6x40x100 tokens) + 180x180 BEV, "
                                "900 queries, 6 layers, d=256 (BASELINE configs[2]/[4])"),
    "coop_lidar": ("CmtLidarHeadCoop", 180, 0, "CMTCoop-L TUMTraf cooperative LiDAR-only: 2 nodes x 180x180 BEV "
                                               "(BASELINE configs[1])"),
    "coop_fusion": ("CmtHeadCoop", 180, 0, "CMTCoop TUMTraf cooperative multimodal: vehicle 1 cam + infrastructure "
                                           "3 cams + 2 LiDAR BEV maps (BASELINE configs[3])"),
    "lidar128": ("CmtLidarHead", 128, 0, "CMT-L LiDAR-only 128x128 BEV (BASELINE configs[0])"),
}


# dram__bytes_read.sum + dram__bytes_write.sum of tc_attn_db_kernel per launch from the committed ncu --set full
# capture (profiles/r1_attn_static_ncu_raw.csv, the static-shift instantiation the bench runs: 1.3950 GB + 0.0118 GB at B=8),
# per frame.  K+V of one layer are
# 57.8 MB per frame; the 3 query blocks of a (frame, head) stream them at different times, hence ~3x.
NCU_DRAM_BYTES_PER_FRAME = {"nusc": (1.394999e9 + 0.011791e9) / 8}


def build_case(workload, B, seed=0):
    kind, bev_hw, n_views, _ = WORKLOADS[workload]
    cfg = synth.head_cfg(kind, num_query=900, num_layers=6, grid=8 * bev_hw)
    cfg["_apply_shared_conv"] = False
    # hot-path scope: the BEV input is the post-shared_conv map (256 channels)
    inputs = synth.make_inputs(kind, B=B, bev_hw=bev_hw, n_views=max(n_views, 1), img_hw=(40, 100),
                               in_channels=256, seed=seed)
    return kind, cfg, inputs


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled every 100 ms during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port: the verbatim reference needs mmcv/mmdet,
    absent on the box), fp32, nn.MultiheadAttention-style attention, all host threads, one frame per step."""
    if rank != 0:
        return
    from oracle import cmt_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    kind, cfg, inputs = build_case(args.workload, 1)
    from cmtcoop_b200.plugin import build_head
    head = build_head({k: v for k, v in cfg.items() if not k.startswith("_")})
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in
          synth.synth_state_dict({k: tuple(v.shape) for k, v in head.state_dict().items()}).items()}
    del head
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.head_forward(sd, cfg, inputs)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    per = float(np.mean(times))
    val = 1.0 / per
    line = dict(metric=METRIC, value=val, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=per * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=WORKLOADS[args.workload][3], frames_per_step=1, precision="fp32",
                            implementation="oracle port of the reference CPU path (nn.MultiheadAttention math)"),
                cpu_baseline=dict(value=val, unit="frames/s", cores=torch.get_num_threads(), kind="port",
                                  sample="1 frame per step"),
                e2e=dict(value=val, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(workload):
    """Bounded CPU sample for the default line: ONE frame of the same workload through the oracle."""
    from oracle import cmt_oracle as O
    from cmtcoop_b200.plugin import build_head
    torch.set_num_threads(os.cpu_count() or 1)
    kind, cfg, inputs = build_case(workload, 1)
    head = build_head({k: v for k, v in cfg.items() if not k.startswith("_")})
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in
          synth.synth_state_dict({k: tuple(v.shape) for k, v in head.state_dict().items()}).items()}
    del head
    iters = 6   # ~10 s of host work on the box's cores
    with torch.no_grad():
        O.head_forward(sd, cfg, inputs)   # untimed warm-up (thread pool, allocator)
        t0 = time.perf_counter()
        for _ in range(iters):
            O.head_forward(sd, cfg, inputs)
        dt = (time.perf_counter() - t0) / iters
    return dict(value=1.0 / dt, unit="frames/s", cores=torch.get_num_threads(), kind="port",
                sample=f"1 frame of the same workload per iteration, 1 warm-up + {iters} timed iterations "
                       f"({dt:.2f} s each), fp32, all host threads")


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="nusc", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=8, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--kv-split", action="store_true",
                    help="BASELINE configs[4] variant: every rank sees the SAME frames and attends 1/N of the K/V tokens; "
                         "one NCCL all-gather of (O, LSE) per decoder layer + log-sum-exp merge (strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5  # bounded: each step is ~10 s of host work
        args.warmup = min(args.warmup, 1)
        run_reference(args, rank, world)
        return

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    import torch.distributed as dist
    from cmtcoop_b200 import ops
    from cmtcoop_b200.plugin import build_head

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    kv_split = args.kv_split and world > 1
    if kv_split:
        args.no_cuda_graph = True   # the per-layer NCCL all-gather stays outside graph capture
    kind, cfg, inputs = build_case(args.workload, B, seed=0 if kv_split else rank)
    head = build_head({k: v for k, v in cfg.items() if not k.startswith("_")})
    synth.load_synth_weights(head, 0)
    head = head.to(dev).eval().set_precision("bf16")
    head.apply_shared_conv = False
    if kv_split:
        head.transformer.enable_kv_split()
    coop = kind.endswith("Coop")
    feat_keys = [k for k, v in inputs.items() if isinstance(v, np.ndarray)]
    host = {k: torch.from_numpy(inputs[k]).pin_memory() for k in feat_keys}
    resident = {k: v.to(dev) for k, v in host.items()}
    metas = inputs["img_metas"]

    def forward(feats):
        g = feats.get
        if coop:
            return head.forward_single(g("vehicle_pts_feats"), g("infrastructure_pts_feats"), g("vehicle_img_feats"),
                                       g("infrastructure_img_feats"), metas)
        return head.forward_single(g("pts_feats"), g("img_feats"), metas)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between a barrier+sync on both sides, CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        # ---- device-resident throughput ----
        for _ in range(args.warmup):
            forward(resident)
        sampler = ClockSampler(local_rank)
        sampler.start()
        n0 = ops.launch_count()
        # per-CTA clock64 totals of the attention kernel (one 8-byte store per CTA): with the event time of the same launch
        # they give the SM clock the kernel actually ran at -- NVML keeps reporting the maximum clock while the kernel
        # runs power-limited, and the MUFU-floor analysis in DESIGN.md is in cycles
        import ctypes
        from cmtcoop_b200 import _lib
        lib = _lib.load()
        lib.cmt_debug_attn_timing.argtypes = [ctypes.c_void_p]
        tbuf = torch.zeros(3 * 96 * 16 + 148, dtype=torch.int64, device=dev)
        lib.cmt_debug_attn_timing(ctypes.c_void_p(tbuf.data_ptr()))
        ops.profile_events("cross_attn", True)
        ms_eager = timed(lambda: forward(resident), args.steps)
        attn_ms = ops.profile_events("cross_attn", False)
        lib.cmt_debug_attn_timing(ctypes.c_void_p(0))
        cyc = float(tbuf[3 * 96 * 16:].max().item())   # the last cross-attention launch of the timed region
        sm_clock_ghz = cyc / (attn_ms[-1] * 1e-3) / 1e9 if attn_ms and cyc > 0 else None
        launches = ops.launch_count() - n0
        ms = ms_eager
        if not args.no_cuda_graph:
            # same forward, same kernels, replayed from a CUDA graph (public API: cmtcoop_b200.runtime.GraphedForward):
            # the eager run above keeps the per-launch attention timings for the roofline
            from cmtcoop_b200.runtime import GraphedForward
            graphed = GraphedForward(head, metas, resident, adopt_inputs=True)
            for _ in range(3):
                graphed()
            ms = timed(lambda: graphed(), args.steps)
        clocks = sampler.stop()

        # ---- end to end: pinned host inputs -> H2D -> forward -> D2H of every task-head tensor ----
        # public serving API: cmtcoop_b200.runtime.PipelinedRunner double-buffers the H2D of step i+1 and
        # the D2H of step i-1 under the compute of step i; every byte still moves inside the timed region
        from cmtcoop_b200.runtime import PipelinedRunner
        runner = PipelinedRunner(head, metas, host, dev, use_cuda_graph=not args.no_cuda_graph)
        runner.run([host] * 3)
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        runner.run([host] * args.steps)
        e1.record()
        barrier()
        ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        ms_e2e = float(ms_t.item())
        h2d, d2h = runner.h2d_bytes, runner.d2h_bytes

    frames = B * (1 if kv_split else world) * args.steps
    value = frames / (ms * 1e-3)
    pk = peaks()
    # roofline of the dominant kernel (tc_attn_kernel): algorithmic flops 4*Nq*N_kv*C per frame per layer
    def node_tokens(prefix):
        n = 0
        p, i = inputs.get(prefix + "pts_feats"), inputs.get(prefix + "img_feats")
        if p is not None:
            n += p.shape[2] * p.shape[3]
        if i is not None:
            n += (i.shape[0] // B) * i.shape[2] * i.shape[3]
        return n

    kv_per_node = [node_tokens("vehicle_"), node_tokens("infrastructure_")] if coop else [node_tokens("")]
    N_kv = float(np.mean(kv_per_node))  # one attention launch per node per layer
    flops_per_launch = 4.0 * 900 * N_kv * 256 * B
    attn_avg_ms = float(np.mean(attn_ms)) if attn_ms else None
    roof = None
    if attn_avg_ms:
        achieved = flops_per_launch / (attn_avg_ms * 1e-3) / 1e12
        roof = dict(bound="tensor", kernel="tc_attn_db_kernel<static shift> (+ online-kernel early exit + merge)", achieved=achieved,
                    peak=pk["bf16_tflops_sustained"], unit="TFLOP/s", frac=achieved / pk["bf16_tflops_sustained"],
                    traffic=NCU_DRAM_BYTES_PER_FRAME.get(args.workload, 0) * B or None, peak_source=pk["source"] + " sustained bf16 (kernel timed inside the step)",
                    launches_timed=len(attn_ms), avg_launch_ms=attn_avg_ms,
                    sm_clock_ghz_under_kernel=sm_clock_ghz,
                    share_of_step=attn_avg_ms * len(attn_ms) / ms_eager,
                    algorithmic_flops_per_launch=flops_per_launch)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(args.workload)
        line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong" if kv_split else "weak",
                    vs_baseline=None,
                    dtype="bf16", data="synthetic",
                    config=dict(workload=WORKLOADS[args.workload][3], frames_per_gpu=B,
                                global_batch=B if kv_split else B * world,
                                parallelism=(f"K/V tokens split x{world}: queries replicated, one NCCL all-gather of "
                                             "(O, LSE) per decoder layer + LSE merge" if kv_split else
                                             f"frame sharding x{world}, no data-path collective"),
                                l2="inputs (%.0f MB per step) exceed the 126 MB L2" % (h2d / 1e6),
                                scope="forward_single from post-shared_conv BEV map + image features to task-head outputs",
                                cuda_graph=not args.no_cuda_graph, eager_ms_per_step=ms_eager / args.steps),
                    clocks=clocks, gpu_launches=launches,
                    e2e=dict(value=frames / (ms_e2e * 1e-3), unit="frames/s", h2d_bytes_per_step=h2d,
                             d2h_bytes_per_step=d2h, ms_per_step=ms_e2e / args.steps),
                    roofline=roof, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
