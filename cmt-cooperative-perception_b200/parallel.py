"""Host-side multi-GPU logic (one process per GPU, torch.distributed; NCCL on the box, gloo in CPU tests).

* Frame sharding: frames are independent on the whole path, so ranks take contiguous blocks of the
  batch and no data-path collective is needed (coop: a frame's two nodes stay on one rank so the
  max-merge is local).
* KV-token split (largest token counts): rank g attends tokens [lo_g, hi_g) only and produces a
  normalised partial O_g with LSE_g; one all-gather of (O_g, LSE_g) per decoder layer, then
  O = sum_g exp(LSE_g - LSE) O_g, LSE = logsumexp_g LSE_g (cmt_lse_merge on the GPU).
"""
from __future__ import annotations

import os

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


def token_rows_copy_plan(bev_shape, img_shape, B: int, lo: int, hi: int, halo: int = 0):
    """Which parts of the NCHW feature maps does a rank of the KV-token split read?  Tokens are the concatenation of
    the BEV positions (y * W + x) and the image positions (v * h * w + y * w + x) of a frame; the rank owns [lo, hi).
    Returns {"pts": [...], "img": [...]}: rectangles (offset, pitch, width, height) in ELEMENTS of the contiguous
    [B, C, H, W] / [B*V, C, h, w] tensors -- `height` runs of `width` elements, `pitch` apart, starting at `offset` --
    that cover every map row holding one of the rank's tokens (whole rows; `halo` extra BEV rows on both sides for the
    3x3 shared_conv).  bev_shape / img_shape: tensor shapes or None."""
    plan = {"pts": [], "img": []}
    n_bev = 0
    if bev_shape is not None:
        Bb, C, H, W = bev_shape
        assert Bb == B
        n_bev = H * W
        a, b = min(lo, n_bev), min(hi, n_bev)
        if b > a:
            y0, y1 = max(0, a // W - halo), min(H, (b + W - 1) // W + halo)
            plan["pts"].append((y0 * W, H * W, (y1 - y0) * W, B * C))
    if img_shape is not None:
        BV, C, h, w = img_shape
        V = BV // B
        assert V * B == BV
        a, b = max(lo, n_bev) - n_bev, min(max(hi, n_bev) - n_bev, V * h * w)
        if b > a:
            va, vb = a // (h * w), (b - 1) // (h * w)
            for v in range(va, vb + 1):
                r0 = (a - v * h * w) // w if v == va else 0
                r1 = (b - v * h * w + w - 1) // w if v == vb else h
                for f in range(B):
                    plan["img"].append((((f * V + v) * C) * h * w + r0 * w, h * w, (r1 - r0) * w, C))
    return plan


def peer_layout(B: int, Nq: int, H: int, n_slots: int, head_dim: int = 32, ctrl_words: int = 64):
    """Word (4-byte) offsets inside one rank's peer-mapped allocation: `n_slots` packed (O | LSE) records, `n_slots`
    context buffers (room for fp32), then the control words {record counters [8], context counters [8], exchange number,
    finished blocks, ...}.  Every record and context buffer starts on a 16-byte boundary (the kernel moves 128-bit
    words)."""
    n_o, n_l = B * Nq * H * head_dim, B * H * Nq
    n_rec = n_o + n_l
    if n_rec % 4 or n_o % 4:
        raise ValueError("peer exchange: B*Nq*H must be a multiple of 4 (16-byte aligned records)")
    ctx_off = n_slots * n_rec
    ctrl_off = ctx_off + n_slots * n_o
    return dict(n_o=n_o, n_l=n_l, n_rec=n_rec, ctx_off=ctx_off, ctrl_off=ctrl_off, total=ctrl_off + ctrl_words,
                state_word=ctrl_off + 16)


class PeerExchange:
    """Peer-mapped buffers for the fused exchange + merge of the KV-token split (ops.lse_merge_peer), one allocation per
    rank from torch's symmetric memory, every rank's mapped into every process of the group over NVLink:
    `n_slots` packed (O | LSE) records of B*Nq*H*32 + B*H*Nq fp32, `n_slots` context buffers of B*Nq*H*32 (room for
    fp32), then 64 control words {record counters [8], context counters [8], exchange number, finished blocks, ...}.
    Nothing here touches the data path: `record(i)` hands out views for the attention kernel to write, `merge(i)`
    launches the one kernel that announces, waits, merges this rank's share of the rows from all ranks' records and
    stores it into all ranks' context buffers.  Consecutive exchanges must use different slots (n_slots >= 2 and
    slot = decoder layer does that, also across forwards and CUDA-graph replays); the returned context is a view of
    slot i's buffer, valid until the next exchange on that slot."""

    CTRL_WORDS = 64

    def __init__(self, group, device, B: int, Nq: int, H: int, n_slots: int):
        import torch.distributed._symmetric_memory as symm
        assert n_slots >= 2, "two consecutive exchanges must not share a record slot"
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        assert self.world <= 8
        self.device = torch.device(device)
        self.shape = (B, Nq, H)
        lay = peer_layout(B, Nq, H, n_slots, ctrl_words=self.CTRL_WORDS)
        self.n_o, self.n_l, self.n_rec = lay["n_o"], lay["n_l"], lay["n_rec"]
        self.n_slots = n_slots
        self.scatter = int(os.environ.get("CMT_PEER_SCATTER", "-1"))   # debugging: force one of the two kernel modes (same on every rank)
        self.ctx_off, self.ctrl_off = lay["ctx_off"], lay["ctrl_off"]   # in fp32 words
        self.buf = symm.empty(lay["total"], dtype=torch.float32, device=self.device)
        self.buf.zero_()
        torch.cuda.current_stream(self.device).synchronize()
        self.handle = symm.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(self.ptrs) == self.world and self.ptrs[self.rank] == self.buf.data_ptr()
        self.arrive_ptrs = [p + self.ctrl_off * 4 for p in self.ptrs]
        self.state_ptr = self.ptrs[self.rank] + (self.ctrl_off + 16) * 4
        dist.barrier(self.group)   # every rank's counters are zero before anyone announces

    def matches(self, B, Nq, H, n_slots):
        return self.shape == (B, Nq, H) and self.n_slots >= n_slots

    def record(self, slot: int):
        """Views (record, o [B,Nq,H*32], lse [B,H,Nq]) of this rank's slot."""
        B, Nq, H = self.shape
        rec = self.buf[slot * self.n_rec:(slot + 1) * self.n_rec]
        return rec, rec[:self.n_o].view(B, Nq, H * 32), rec[self.n_o:].view(B, H, Nq)

    def merge(self, slot: int, o_dtype=torch.bfloat16):
        from . import ops
        B, Nq, H = self.shape
        roff, coff = slot * self.n_rec * 4, (self.ctx_off + slot * self.n_o) * 4
        ops.lse_merge_peer([p + roff for p in self.ptrs], [p + coff for p in self.ptrs], self.arrive_ptrs, self.state_ptr,
                           self.rank, B, Nq, H, self.device, o_dtype, self.scatter)
        ctx = self.buf[self.ctx_off + slot * self.n_o:self.ctx_off + (slot + 1) * self.n_o]
        if o_dtype == torch.float32:
            return ctx.view(B, Nq, H * 32)
        return ctx.view(o_dtype)[:self.n_o].view(B, Nq, H * 32)


def merge_partials_reference(o_all: torch.Tensor, l_all: torch.Tensor, num_heads: int):
    """Plain-torch statement of the LSE merge (what cmt_lse_merge computes); used by CPU tests."""
    G, B, Nq, C = o_all.shape
    lse = torch.logsumexp(l_all, dim=0)                                   # [B,H,Nq]
    w = torch.exp(l_all - lse)                                            # [G,B,H,Nq]
    w = torch.nan_to_num(w, nan=0.0)
    o = (o_all.view(G, B, Nq, num_heads, C // num_heads) * w.permute(0, 1, 3, 2).unsqueeze(-1)).sum(0)
    return o.reshape(B, Nq, C), lse
