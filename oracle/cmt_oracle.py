"""TEST INFRASTRUCTURE ONLY -- the CPU oracle of the CMT / CMTCoop token-fusion hot path.

A plain fp32 (optionally fp64) torch-CPU restatement of the reference algorithm, function by
function, each citing the reference file:line it follows (paths relative to
/root/reference/projects/mmdet3d_plugin/).  It takes a flat state dict (same keys as the
reference modules) and numpy/torch inputs; it never touches CUDA and never imports the product
package's kernels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module, and only as the checker -- never as the thing shipped.

Pinning: the reference ships no tests or golden vectors, so this restatement
is pinned against the *verbatim reference code* run on CPU in the build container
(oracle/ref_stub.py + oracle/make_golden.py -> tests/golden/*.npz, checked by
tests/test_oracle_golden.py everywhere and by tests/test_oracle_vs_reference.py where
/root/reference exists).  The third-party pieces the reference calls (mmcv BaseTransformerLayer /
FFN / MultiheadAttention wrapper, mmdet inverse_sigmoid) are restated from mmcv-full 1.6.2 /
mmdet 2.28.2 and are "parity unpinned" by any reference-owned test.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _t(x, dtype=torch.float32):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(dtype)


# ---------------------------------------------------------------------------------------------
# position encodings
# ---------------------------------------------------------------------------------------------
def pos2embed(pos, num_pos_feats=128):
    """models/dense_heads/cmt_head.py:40-50 (note: `temperature` is unused there)."""
    pos = pos * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = 2 * (dim_t // 2) / num_pos_feats + 1
    dim_t = dim_t.to(pos.dtype)
    px = pos[..., 0, None] / dim_t
    py = pos[..., 1, None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=-1).flatten(-2)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=-1).flatten(-2)
    return torch.cat((py, px), dim=-1)


def coords_bev(grid_size, downsample_scale=8):
    """cmt_head.py:324-337: BEV cell centres, token t = i*W + j <-> ((j+.5)/W, (i+.5)/H)."""
    x_size = grid_size[1] // downsample_scale
    y_size = grid_size[0] // downsample_scale
    by, bx = torch.meshgrid(torch.linspace(0, x_size - 1, x_size), torch.linspace(0, y_size - 1, y_size),
                            indexing="ij")
    bx = (bx + 0.5) / x_size
    by = (by + 0.5) / y_size
    return torch.cat([bx[None], by[None]], dim=0).view(2, -1).transpose(1, 0)


def img2lidar_f32(lidar2img_list):
    """cmt_head.py:428-429: float64 inverse on the host, then cast to fp32. -> [n,4,4] float32 numpy."""
    return np.stack([np.linalg.inv(np.asarray(m, dtype=np.float64)) for m in lidar2img_list]).astype(np.float32)


def depth_bins(depth_num, pc_range, dtype=torch.float32):
    """cmt_head.py:422."""
    return 1 + torch.arange(depth_num).to(dtype) * (pc_range[3] - 1) / depth_num


def ray_coords(img2lidar, H, W, depth_num, pad_h, pad_w, pc_range, dtype=torch.float32):
    """cmt_head.py:417-432: normalised lifted points, [n_cam,H,W,depth_num*3] (feature 3k+c)."""
    M = _t(img2lidar, dtype)
    ch = torch.arange(H).to(dtype) * pad_h / H
    cw = torch.arange(W).to(dtype) * pad_w / W
    cd = depth_bins(depth_num, pc_range, dtype)
    gh, gw, gd = torch.meshgrid([ch, cw, cd], indexing="ij")
    coords = torch.stack([gw, gh, gd, torch.ones_like(gh)], dim=-1)
    coords[..., :2] = coords[..., :2] * coords[..., 2:3]
    c3 = torch.einsum("hwdo,bco->bhwdc", coords, M)
    lo = torch.tensor(pc_range[:3], dtype=dtype)
    hi = torch.tensor(pc_range[3:], dtype=dtype)
    c3 = (c3[..., :3] - lo) / (hi - lo)
    return c3.reshape(*c3.shape[:-2], -1)


def mlp2(x, sd, prefix):
    """nn.Sequential(Linear, ReLU, Linear): cmt_head.py:292-301."""
    h = F.relu(F.linear(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"]))
    return F.linear(h, sd[prefix + ".2.weight"], sd[prefix + ".2.bias"])


def inverse_sigmoid(x, eps=1e-5):
    """mmdet 2.28.2 models/utils/transformer.py inverse_sigmoid (third-party, restated)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def rv_query_feats(ref, lidar2img, img2lidar, depth_num, pad_h, pad_w, pc_range):
    """cmt_head.py:439-464 up to (not including) rv_embedding.
    ref [B,Nq,3] in [0,1]; lidar2img/img2lidar [B,V,4,4] -> feats [B,V,Nq,depth_num*3], mask [B,V,Nq] bool."""
    dtype = ref.dtype
    L = _t(lidar2img, dtype)
    M = _t(img2lidar, dtype)
    lo = torch.tensor(pc_range[:3], dtype=dtype)
    hi = torch.tensor(pc_range[3:], dtype=dtype)
    P = ref * (hi - lo) + lo
    proj = torch.einsum("bnd,bvcd->bvnc", torch.cat([P, torch.ones_like(P[..., :1])], dim=-1), L)
    pc = proj.clone()
    zmask = pc[..., 2:3] > 0
    pc[..., :3] = proj[..., :3] / (proj[..., 2:3] + zmask * 1e-6 - (~zmask) * 1e-6)
    mask = (pc[..., 0] < pad_w) & (pc[..., 0] >= 0) & (pc[..., 1] < pad_h) & (pc[..., 1] >= 0)
    mask &= zmask.squeeze(-1)
    cd = depth_bins(depth_num, pc_range, dtype)
    pc = torch.einsum("bvnc,d->bvndc", pc, cd)
    pc = torch.cat([pc[..., :3], torch.ones_like(pc[..., :1])], dim=-1)
    back = torch.einsum("bvndo,bvco->bvndc", pc, M)
    back = (back[..., :3] - lo) / (hi - lo)
    return back.reshape(*back.shape[:-2], -1), mask


# ---------------------------------------------------------------------------------------------
# attention / decoder
# ---------------------------------------------------------------------------------------------
def mha(query, key, value, sd, prefix, num_heads=8, head_chunk=2, key_keep=None):
    """nn.MultiheadAttention forward (the CPU/fp32 cross-attention oracle path,
    models/utils/petr_transformer.py:37-177 wraps it; same parameter names as FlashMHA,
    models/utils/attention.py:95-138).  query [Nq,B,C], key/value [Nk,B,C] seq-first.
    key_keep [B,Nk] bool: the key_padding_mask branch of FlashAttention.forward (attention.py:76-90).  There the
    mask goes to flash_attn.bert_padding.unpad_input (flash-attn 0.2.2, not under /root/reference: it keeps the
    entries where the mask is True -- `indices = nonzero(mask.flatten())` -- and builds cu_seqlens_k from the
    per-frame counts), so a False key simply does not exist for its frame: weight zero here."""
    w = sd[prefix + ".in_proj_weight"]
    b = sd[prefix + ".in_proj_bias"]
    C = query.shape[-1]
    d = C // num_heads
    q = F.linear(query, w[:C], b[:C])
    k = F.linear(key, w[C:2 * C], b[C:2 * C])
    v = F.linear(value, w[2 * C:], b[2 * C:])
    Nq, B, _ = q.shape
    Nk = k.shape[0]
    q = q.reshape(Nq, B, num_heads, d).permute(1, 2, 0, 3)  # B,H,Nq,d
    k = k.reshape(Nk, B, num_heads, d).permute(1, 2, 0, 3)
    v = v.reshape(Nk, B, num_heads, d).permute(1, 2, 0, 3)
    out = torch.empty_like(q)
    scale = 1.0 / math.sqrt(d)
    for h0 in range(0, num_heads, head_chunk):  # bounded memory for 900 x 56400 score maps
        s = torch.matmul(q[:, h0:h0 + head_chunk] * scale, k[:, h0:h0 + head_chunk].transpose(-1, -2))
        if key_keep is not None:
            s = s.masked_fill(~key_keep.bool()[:, None, None, :], float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[:, h0:h0 + head_chunk] = torch.matmul(p, v[:, h0:h0 + head_chunk])
    out = out.permute(2, 0, 1, 3).reshape(Nq, B, C)
    return F.linear(out, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def decoder(sd, prefix, memory, pos_embed, query_embed, num_layers, num_heads=8):
    """PETRTransformerDecoder (petr_transformer.py:347-371) over PETRTransformerDecoderLayer /
    mmcv BaseTransformerLayer with operation order self_attn, norm, cross_attn, norm, ffn, norm
    (post-norm), eval mode.  memory/pos_embed [Nk,B,C], query_embed [Nq,B,C] -> [L,Nq,B,C]."""
    C = memory.shape[-1]
    x = torch.zeros_like(query_embed)  # cmt_transformer.py:114
    key = memory + pos_embed           # petr_transformer.py:298-299 (layer-invariant)
    outs = []

    def ln(t, p):
        return F.layer_norm(t, (C,), sd[p + ".weight"], sd[p + ".bias"], 1e-5)

    for l in range(num_layers):
        lp = f"{prefix}.layers.{l}"
        qk = x + query_embed
        x = x + mha(qk, qk, x, sd, lp + ".attentions.0.attn", num_heads)          # self-attn, key_pos=query_pos
        x = ln(x, lp + ".norms.0")
        x = x + mha(x + query_embed, key, memory, sd, lp + ".attentions.1.attn", num_heads)
        x = ln(x, lp + ".norms.1")
        h = F.relu(F.linear(x, sd[lp + ".ffns.0.layers.0.0.weight"], sd[lp + ".ffns.0.layers.0.0.bias"]))
        x = x + F.linear(h, sd[lp + ".ffns.0.layers.1.weight"], sd[lp + ".ffns.0.layers.1.bias"])
        x = ln(x, lp + ".norms.2")
        outs.append(ln(x, prefix + ".post_norm"))
    return torch.stack(outs)


def tokens(x_bev, x_img, bev_pos, rv_pos, B):
    """cmt_transformer.py:105-110 (also :186-187, :262-263): memory/pos [N_kv,B,C], BEV tokens first
    (row-major h,w), then image tokens view-major."""
    mems, poss = [], []
    if x_bev is not None:
        mems.append(x_bev.flatten(2).permute(2, 0, 1))
        poss.append(bev_pos.unsqueeze(1).repeat(1, B, 1))
    if x_img is not None:
        BV, C, h, w = x_img.shape
        V = BV // B
        mems.append(x_img.reshape(B, V, C, h * w).permute(1, 3, 0, 2).reshape(V * h * w, B, C))
        poss.append(rv_pos.reshape(B, V, h * w, -1).permute(1, 2, 0, 3).reshape(V * h * w, B, -1))
    return torch.cat(mems, 0), torch.cat(poss, 0)


# ---------------------------------------------------------------------------------------------
# task heads / box decode
# ---------------------------------------------------------------------------------------------
def group_layer_norm_1d(x, weight, bias, groups, eps=1e-6):
    """cmt_head.py:53-94."""
    N, C, L = x.shape
    xg = x.view(N, groups, C // groups, L)
    mu = xg.mean(2, keepdim=True)
    var = (xg - mu).pow(2).mean(2, keepdim=True)
    y = (xg - mu) / (var + eps).sqrt()
    return weight.view(1, C, 1) * y.view(N, C, L) + bias.view(1, C, 1)


def separate_task_head(outs_dec, sd, prefix, head_names, final_kernel):
    """SeparateTaskHead.forward (cmt_head.py:174-203). outs_dec [L,B,Nq,C] -> {name: [L,B,Nq,c_out]}."""
    L, B, Nq, C = outs_dec.shape
    x = outs_dec.permute(1, 0, 3, 2).reshape(B, L * C, Nq)
    ret = {}
    pad = final_kernel // 2
    for name in head_names:
        p = f"{prefix}.{name}"
        y = F.conv1d(x, sd[p + ".0.weight"], None, padding=pad, groups=L)
        y = group_layer_norm_1d(y, sd[p + ".1.weight"], sd[p + ".1.bias"], L)
        y = F.relu(y)
        y = F.conv1d(y, sd[p + ".3.weight"], sd[p + ".3.bias"], padding=pad, groups=L)
        ret[name] = y.view(B, L, -1, Nq).permute(1, 0, 3, 2)
    return ret


def decode_outputs(outs_dec, reference_points, sd, cfg):
    """Tail of forward_single (cmt_head.py:501-513): task heads + reference-point decode."""
    pc = cfg["bbox_coder"]["pc_range"]
    fk = cfg["separate_head"]["final_kernel"]
    reference = inverse_sigmoid(reference_points.clone())
    rets = []
    for t, task in enumerate(cfg["tasks"]):
        names = list(cfg["common_heads"].keys()) + ["cls_logits"]
        outs = separate_task_head(outs_dec, sd, f"task_heads.{t}", names, fk)
        center = (outs["center"] + reference[None, :, :, :2]).sigmoid()
        height = (outs["height"] + reference[None, :, :, 2:3]).sigmoid()
        _c = torch.zeros_like(center)
        _h = torch.zeros_like(height)
        _c[..., 0:1] = center[..., 0:1] * (pc[3] - pc[0]) + pc[0]
        _c[..., 1:2] = center[..., 1:2] * (pc[4] - pc[1]) + pc[1]
        _h[..., 0:1] = height[..., 0:1] * (pc[5] - pc[2]) + pc[2]
        outs["center"], outs["height"] = _c, _h
        rets.append(outs)
    return rets


def bbox_decode(ret_dicts, cfg):
    """MultiTaskBBoxCoder.decode (core/bbox/coders/multi_task_bbox_coder.py:46-141) +
    denormalize_bbox (core/bbox/util.py:37-68) + the z shift of get_bboxes (cmt_head.py:912).
    Returns per frame dict(bboxes, scores, labels, topk_index [int64 flat index before the range mask])."""
    bc = cfg["bbox_coder"]
    num_classes = bc["num_classes"]
    max_num = bc["max_num"]
    post = torch.tensor(bc["post_center_range"], dtype=torch.float32)
    bbox_l, logit_l, tid_l = [], [], []
    for t, d in enumerate(ret_dicts):
        bbox_l.append(torch.cat((d["center"][-1], d["height"][-1], d["dim"][-1], d["rot"][-1], d["vel"][-1]), -1))
        logit_l.append(d["cls_logits"][-1])
        tid_l.append(torch.full(d["cls_logits"][-1].shape, t, dtype=torch.int32))
    logits = torch.cat(logit_l, -1)
    bboxes = torch.cat(bbox_l, 1)
    tids = torch.cat(tid_l, -1)
    out = []
    for i in range(logits.shape[0]):
        nq = logits[i].shape[0]
        scores, idx = logits[i].sigmoid().view(-1).topk(max_num)
        labels = idx % num_classes
        qidx = idx // num_classes
        task_index = torch.gather(tids[i], 1, labels.unsqueeze(1)).squeeze()
        bp = bboxes[i][task_index * nq + qidx]
        cx, cy, cz = bp[..., 0:1], bp[..., 1:2], bp[..., 2:3]
        w, l, h = bp[..., 3:4].exp(), bp[..., 4:5].exp(), bp[..., 5:6].exp()
        rot = torch.atan2(bp[..., 6:7], bp[..., 7:8])
        box = torch.cat([cx, cy, cz, w, l, h, rot, bp[..., 8:9], bp[..., 9:10]], -1)
        m = (box[..., :3] >= post[:3]).all(1) & (box[..., :3] <= post[3:]).all(1)
        box = box[m].clone()
        box[:, 2] = box[:, 2] - box[:, 5] * 0.5
        out.append(dict(bboxes=box, scores=scores[m], labels=labels[m], topk_index=idx))
    return out


# ---------------------------------------------------------------------------------------------
# orchestrators
# ---------------------------------------------------------------------------------------------
def _shared_conv(x, sd):
    """ConvModule 3x3 conv (no bias) + BN2d(eval) + ReLU: cmt_head.py:280-287,481."""
    y = F.conv2d(x, sd["shared_conv.conv.weight"], None, padding=1)
    y = F.batch_norm(y, sd["shared_conv.bn.running_mean"], sd["shared_conv.bn.running_var"],
                     sd["shared_conv.bn.weight"], sd["shared_conv.bn.bias"], False, 0.0, 1e-5)
    return F.relu(y)


def node_outs_dec(sd, cfg, x, x_img, metas, stages=None):
    """get_outs_dec (cmt_head_coop.py:341-360) == the body of CmtHead.forward_single up to
    nan_to_num (cmt_head.py:481-499), for any of the three modality variants.
    x [B,Cin,Hb,Wb] | None, x_img [B*V,C,h,w] | None, metas: list of dicts with lidar2img/pad_shape."""
    hidden = cfg["hidden_dim"]
    depth_num = cfg.get("depth_num", 64)
    pc = cfg["bbox_coder"]["pc_range"]
    nl = cfg["transformer"]["decoder"]["num_layers"]
    B = len(metas)
    ref = sd["reference_points.weight"].unsqueeze(0).repeat(B, 1, 1)      # cmt_head.py:410-411 (eval)
    bev_pos = rv_pos = None
    if x is not None:
        if cfg.get("_apply_shared_conv", True):  # bench scope: "CmtTransformer+PE" starts after shared_conv
            x = _shared_conv(x, sd)
        grid = cfg["test_cfg"]["grid_size"]
        bev_pos = mlp2(pos2embed(coords_bev(grid, cfg.get("downsample_scale", 8)), hidden), sd, "bev_embedding")
    r = inverse_sigmoid(ref.clone()).sigmoid()                             # cmt_head.py:470
    q_embed = mlp2(pos2embed(r, hidden), sd, "bev_embedding")              # :435-437
    if x_img is not None:
        BV, _, h, w = x_img.shape
        pad_h, pad_w, _ = metas[0]["pad_shape"][0]
        l2i = np.stack([np.asarray(m["lidar2img"], dtype=np.float64) for m in metas])          # [B,V,4,4]
        i2l = np.stack([np.linalg.inv(np.asarray(m["lidar2img"], dtype=np.float64)) for m in metas])
        i2l32 = torch.from_numpy(i2l).float()
        l2i32 = torch.from_numpy(l2i).float()
        coords = ray_coords(i2l32.reshape(-1, 4, 4), h, w, depth_num, pad_h, pad_w, pc)
        rv_pos = mlp2(coords, sd, "rv_embedding")                         # [B*V,h,w,C]
        feats, mask = rv_query_feats(r, l2i32, i2l32, depth_num, pad_h, pad_w, pc)
        rv_q = (mlp2(feats, sd, "rv_embedding") * mask.unsqueeze(-1)).sum(dim=1)   # :465-466
        q_embed = q_embed + rv_q                                          # :492
        if stages is not None:
            stages["ray_coords"] = coords
            stages["rv_pos"] = rv_pos
            stages["rv_query_feats"] = feats
            stages["rv_query_mask"] = mask
    memory, pos = tokens(x, x_img, bev_pos, rv_pos, B)
    out = decoder(sd, "transformer.decoder", memory, pos, q_embed.transpose(0, 1), nl)
    out = torch.nan_to_num(out.transpose(1, 2))                           # [L,B,Nq,C]
    if stages is not None:
        stages["bev_pos"] = bev_pos
        stages["query_embed"] = q_embed
        stages["memory"] = memory
    return out


def strip_prefix_metas(metas, prefix, ignore):
    """filter_img_metas (cmt_head_coop.py:41-69)."""
    out = []
    for m in metas:
        f = {}
        for k, v in m.items():
            if k.startswith(prefix):
                f[k[len(prefix):]] = v
            elif not k.startswith(ignore):
                f[k] = v
        f["node"] = prefix
        out.append(f)
    return out


def head_forward(sd, cfg, inputs, stages=None):
    """forward_single of any of the six head classes (cmt_head.py:475-547, :929-999, :1014-1085;
    cmt_head_coop.py:362-437, :839-911, :946-1017).  `inputs` as produced by synth.make_inputs.
    Returns (ret_dicts, outs_dec)."""
    sd = {k: _t(v) if not (isinstance(v, torch.Tensor) and v.dtype == torch.int64) else v for k, v in sd.items()}
    kind = cfg["type"]
    metas = inputs["img_metas"]
    B = len(metas)

    def tt(a):
        return None if a is None else _t(a)

    if not kind.endswith("Coop"):
        outs_dec = node_outs_dec(sd, cfg, tt(inputs.get("pts_feats")), tt(inputs.get("img_feats")), metas, stages)
    else:
        per_node = []
        for node, ign in (("vehicle", "infrastructure_"), ("infrastructure", "vehicle_")):
            x = tt(inputs.get(f"{node}_pts_feats"))
            xi = tt(inputs.get(f"{node}_img_feats"))
            if x is None and xi is None:
                continue
            per_node.append(node_outs_dec(sd, cfg, x, xi, strip_prefix_metas(metas, node + "_", ign),
                                          stages if node == "vehicle" else None))
        outs_dec = per_node[0] if len(per_node) == 1 else torch.max(torch.stack(per_node), 0).values  # :383-389
    ref = sd["reference_points.weight"].unsqueeze(0).repeat(B, 1, 1)
    return decode_outputs(outs_dec, ref, sd, cfg), outs_dec


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def box_set_overlap(a, b, tol=1e-3):
    """Fraction of rows of `a` that have a row of `b` within `tol` (relative L-inf).  Order-free."""
    if a.numel() == 0:
        return 1.0 if b.numel() == 0 else 0.0
    if b.numel() == 0:
        return 0.0
    d = (a[:, None, :].double() - b[None, :, :].double()).abs()
    scale = b.abs().double().amax(dim=0).clamp_min(1.0)
    ok = (d / scale).amax(dim=-1) < tol
    return float(ok.any(dim=1).double().mean())


def topk_tie_tolerant_equal(idx_ours, scores_ours_all, idx_ref, scores_ref_all, tau):
    """every index we select has reference score >= reference k-th score - tau and
    every reference-selected index has our score >= our k-th score - tau."""
    kth_ref = scores_ref_all[idx_ref].min()
    kth_ours = scores_ours_all[idx_ours].min()
    a = bool((scores_ref_all[idx_ours] >= kth_ref - tau).all())
    b = bool((scores_ours_all[idx_ref] >= kth_ours - tau).all())
    return a and b
