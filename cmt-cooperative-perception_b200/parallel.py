"""Host-side multi-GPU logic (one process per GPU, torch.distributed; NCCL on the box, gloo in CPU tests).

* Frame sharding: frames are independent on the whole path, so ranks take contiguous blocks of the
  batch and no data-path collective is needed (coop: a frame's two nodes stay on one rank so the
  max-merge is local).
* KV-token split (largest token counts): rank g attends tokens [lo_g, hi_g) only and produces a
  normalised partial O_g with LSE_g; one all-gather of (O_g, LSE_g) per decoder layer, then
  O = sum_g exp(LSE_g - LSE) O_g, LSE = logsumexp_g LSE_g (cmt_lse_merge on the GPU).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

KV_TILE = 128  # attention kernel tile: ranges are aligned to it so no rank gets a ragged first tile


def shard_frames(n_frames: int, rank: int, world: int):
    """Contiguous block [lo, hi) of frames for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def kv_split_range(n_kv: int, rank: int, world: int, tile: int = KV_TILE):
    """Token range [lo, hi) of rank `rank`: tile-aligned chunks, the last ranks may be short or empty."""
    tiles = (n_kv + tile - 1) // tile
    per = (tiles + world - 1) // world
    lo = min(rank * per * tile, n_kv)
    hi = min((rank + 1) * per * tile, n_kv)
    return lo, hi


def gather_partials(o_part: torch.Tensor, lse_part: torch.Tensor, group=None):
    """All-gather the per-rank attention partials. o_part [B,Nq,C] fp32, lse_part [B,H,Nq] fp32 ->
    ([G,B,Nq,C], [G,B,H,Nq]).  The only collective on the hot path (6 per forward)."""
    world = dist.get_world_size(group)
    # all_gather_into_tensor concatenates along dim 0 (gloo insists on that exact shape): [G*B, ...]
    o_all = torch.empty((world * o_part.shape[0],) + tuple(o_part.shape[1:]), dtype=o_part.dtype, device=o_part.device)
    l_all = torch.empty((world * lse_part.shape[0],) + tuple(lse_part.shape[1:]), dtype=lse_part.dtype,
                        device=lse_part.device)
    dist.all_gather_into_tensor(o_all, o_part.contiguous(), group=group)
    dist.all_gather_into_tensor(l_all, lse_part.contiguous(), group=group)
    return o_all.view((world,) + tuple(o_part.shape)), l_all.view((world,) + tuple(lse_part.shape))


def gather_packed(record: torch.Tensor, group=None):
    """All-gather of one packed (O | LSE) record per rank (ops.packed_partial): flat fp32 [n] -> [G * n].  ONE collective
    per decoder layer is the whole data-path communication of the KV-token split."""
    world = dist.get_world_size(group)
    out = torch.empty((world * record.numel(),), dtype=record.dtype, device=record.device)
    dist.all_gather_into_tensor(out, record, group=group)
    return out


def merge_partials_reference(o_all: torch.Tensor, l_all: torch.Tensor, num_heads: int):
    """Plain-torch statement of the LSE merge (what cmt_lse_merge computes); used by CPU tests."""
    G, B, Nq, C = o_all.shape
    lse = torch.logsumexp(l_all, dim=0)                                   # [B,H,Nq]
    w = torch.exp(l_all - lse)                                            # [G,B,H,Nq]
    w = torch.nan_to_num(w, nan=0.0)
    o = (o_all.view(G, B, Nq, num_heads, C // num_heads) * w.permute(0, 1, 3, 2).unsqueeze(-1)).sum(0)
    return o.reshape(B, Nq, C), lse
