"""world_size-2 runs of the multi-GPU host logic on CPU (gloo): frame sharding needs no collective and
covers the batch; the KV-token split + all-gather + LSE merge reproduces full attention."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmtcoop_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial_attention(q, k, v, lo, hi):
    """softmax over tokens [lo,hi) only, normalised, plus natural-log LSE (what cmt_cross_attn_fwd returns)."""
    B, Nq, C = q.shape
    H = 8
    qh = q.view(B, Nq, H, C // H).permute(0, 2, 1, 3)
    kh = k[:, lo:hi].reshape(B, hi - lo, H, C // H).permute(0, 2, 1, 3)
    vh = v[:, lo:hi].reshape(B, hi - lo, H, C // H).permute(0, 2, 1, 3)
    if hi <= lo:
        return torch.zeros(B, Nq, C), torch.full((B, H, Nq), float("-inf"))
    s = qh @ kh.transpose(-1, -2) / math.sqrt(C // H)
    return (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, Nq, C), torch.logsumexp(s, -1)


def _worker(rank, world, port, n_kv, q_out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)          # replicated queries / tokens on every rank
        q = torch.randn(2, 50, 256, generator=g)
        k = torch.randn(2, n_kv, 256, generator=g)
        v = torch.randn(2, n_kv, 256, generator=g)
        lo, hi = parallel.kv_split_range(n_kv, rank, world)
        o_p, l_p = _partial_attention(q, k, v, lo, hi)
        o_all, l_all = parallel.gather_partials(o_p, l_p)
        o, lse = parallel.merge_partials_reference(o_all, l_all, 8)
        want_o, want_l = _partial_attention(q, k, v, 0, n_kv)
        ok = torch.allclose(o, want_o, atol=1e-5) and torch.allclose(lse, want_l, atol=1e-5)
        # the packed (O | LSE) record: ONE all-gather per decoder layer carries both (what the GPU path sends)
        rec = torch.cat([o_p.reshape(-1), l_p.reshape(-1)])
        allrec = parallel.gather_packed(rec).view(world, -1)
        n_o = o_p.numel()
        o2, lse2 = parallel.merge_partials_reference(allrec[:, :n_o].reshape(world, *o_p.shape),
                                                     allrec[:, n_o:].reshape(world, *l_p.shape), 8)
        ok = ok and torch.equal(o2, o) and torch.equal(lse2, lse)
        # frame sharding: disjoint contiguous blocks, no collective needed to compute them
        flo, fhi = parallel.shard_frames(13, rank, world)
        mine = torch.zeros(13)
        mine[flo:fhi] = 1
        dist.all_reduce(mine)
        ok = ok and bool((mine == 1).all())
        q_out.put((rank, ok, (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_kv", [300, 129, 100])
def test_kv_split_merge_world2(n_kv):
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_kv, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q_out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
    ranges = dict((r, rg) for r, _, rg in results)
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == n_kv
