"""The CPU oracle (oracle/cmt_oracle.py) against the committed golden vectors, which are outputs
of the UNMODIFIED reference code run on CPU (oracle/make_golden.py).  fp32 on both sides; the only
differences are summation order inside BLAS calls, so the tolerance is a few fp32 ulps of rel-L2."""
import os

import numpy as np
import pytest
import torch

from cmtcoop_b200 import synth
from oracle import cmt_oracle as O

TOL = 2e-5  # rel-L2, fp32 vs fp32 with different op fusion / summation order


def _state_dict_shapes(kind, cfg):
    from cmtcoop_b200.plugin import build_head
    head = build_head(cfg)
    return {k: tuple(v.shape) for k, v in head.state_dict().items()}


@pytest.mark.parametrize("kind", synth.HEAD_KINDS)
def test_oracle_matches_reference_golden(kind, golden_dir):
    gold = np.load(os.path.join(golden_dir, f"{kind}_mini.npz"))
    cfg, inputs = synth.mini_case(kind)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.synth_state_dict(_state_dict_shapes(kind, cfg)).items()}
    stages = {}
    rets, outs_dec = O.head_forward(sd, cfg, inputs, stages)
    if "outs_dec" in gold:
        assert O.rel_l2(outs_dec, torch.from_numpy(gold["outs_dec"])) < TOL
    for name in ("center", "height", "dim", "rot", "vel", "cls_logits"):
        g = torch.from_numpy(gold[f"task0.{name}"])
        assert rets[0][name].shape == g.shape
        assert O.rel_l2(rets[0][name], g) < TOL, name
    if kind == "CmtHead":
        assert O.rel_l2(stages["ray_coords"], torch.from_numpy(gold["ray_coords"])) < 1e-6
        assert O.rel_l2(stages["rv_pos"], torch.from_numpy(gold["rv_pos"])) < TOL
        assert O.rel_l2(stages["rv_query_feats"], torch.from_numpy(gold["rv_query_feats"])) < 1e-5
        assert O.rel_l2(stages["bev_pos"], torch.from_numpy(gold["bev_pos"])) < TOL
        sincos = O.pos2embed(O.coords_bev(cfg["test_cfg"]["grid_size"]), cfg["hidden_dim"])
        assert torch.equal(sincos, torch.from_numpy(gold["bev_sincos"]))  # bit-exact PE token indexing
    boxes = O.bbox_decode(rets, cfg)
    for i, b in enumerate(boxes):
        g_scores = torch.from_numpy(gold[f"boxes{i}.scores"])
        assert b["scores"].shape == g_scores.shape
        assert torch.allclose(b["scores"], g_scores, atol=1e-5)
        # top-k over near-tied scores of an untrained decoder is order-unstable:
        # compare the box SETS, tolerating a few swaps at the k-th score boundary.
        assert O.box_set_overlap(b["bboxes"], torch.from_numpy(gold[f"boxes{i}.bboxes"]), 1e-3) >= 0.9


def test_oracle_mha_key_padding_equals_unpadded_keys():
    """attention.py:76-90: with a key_padding_mask the reference gathers the kept keys of every frame
    (unpad_input) and attends over the packed sequence.  The oracle's masked form must equal attention over the
    explicitly gathered keys, frame by frame."""
    import torch
    from oracle import cmt_oracle as O
    g = torch.Generator().manual_seed(5)
    C, H, Nq, Nk, B = 64, 8, 7, 37, 3
    sd = {"a.in_proj_weight": torch.randn(3 * C, C, generator=g) * 0.2, "a.in_proj_bias": torch.randn(3 * C, generator=g) * 0.1,
          "a.out_proj.weight": torch.randn(C, C, generator=g) * 0.2, "a.out_proj.bias": torch.randn(C, generator=g) * 0.1}
    q = torch.randn(Nq, B, C, generator=g)
    k = torch.randn(Nk, B, C, generator=g)
    v = torch.randn(Nk, B, C, generator=g)
    keep = torch.rand(B, Nk, generator=g) > 0.4
    keep[1, :20] = False           # a long padded prefix
    keep[2] = True                 # one frame without padding
    got = O.mha(q, k, v, sd, "a", num_heads=H, key_keep=keep)
    for b in range(B):
        idx = keep[b].nonzero().flatten()
        want = O.mha(q[:, b:b + 1], k[idx][:, b:b + 1], v[idx][:, b:b + 1], sd, "a", num_heads=H)
        assert torch.allclose(got[:, b:b + 1], want, atol=1e-5, rtol=1e-5)
