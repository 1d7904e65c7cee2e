"""Deterministic synthetic workloads: reference-shaped head configs, weights and inputs.

Everything is generated from ``numpy.random.RandomState`` (MT19937, stable across numpy versions
and machines) so that this container (golden-vector generation from the verbatim reference) and the
GPU box (parity tests, bench) see bit-identical weights and inputs without shipping large files.
Shapes and config dicts follow the reference configs:
  projects/configs/CMT_Nuscenes/fusion/cmt_voxel0075_vov_1600x640_cbgs.py:237-293 (CmtHead)
  projects/configs/CMT_Nuscenes/lidar/cmt_lidar_voxel0075_cbgs.py:199-247        (CmtLidarHead)
  projects/configs/CMTCoop_TUMTraf/*/coop/*.py                                    (coop heads)
"""
from __future__ import annotations

import copy
import zlib

import numpy as np

NUSC_CLASSES = ['car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'barrier', 'motorcycle',
                'bicycle', 'pedestrian', 'traffic_cone']
TUMTRAF_CLASSES = ['CAR', 'TRAILER', 'TRUCK', 'VAN', 'PEDESTRIAN', 'BUS', 'BICYCLE']
NUSC_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
TUMTRAF_RANGE = [-72.0, -72.0, -8.0, 72.0, 72.0, 0.0]

HEAD_KINDS = ('CmtHead', 'CmtLidarHead', 'CmtImageHead', 'CmtHeadCoop', 'CmtLidarHeadCoop', 'CmtImageHeadCoop')
_TRANSFORMER_OF = {
    'CmtHead': 'CmtTransformer', 'CmtHeadCoop': 'CmtTransformer',
    'CmtLidarHead': 'CmtLidarTransformer', 'CmtLidarHeadCoop': 'CmtLidarTransformer',
    'CmtImageHead': 'CmtImageTransformer', 'CmtImageHeadCoop': 'CmtImageTransformer',
}


def head_cfg(kind='CmtHead', *, num_query=900, num_layers=6, grid=1440, pc_range=None, classes=None,
             final_kernel=None, hidden_dim=256, in_channels=512, max_num=300, cross_attn='PETRMultiheadFlashAttention'):
    """A `pts_bbox_head` dict shaped like the reference configs' (same keys, same nesting)."""
    assert kind in HEAD_KINDS
    coop = kind.endswith('Coop')
    pc_range = list(pc_range if pc_range is not None else (TUMTRAF_RANGE if coop else NUSC_RANGE))
    classes = list(classes if classes is not None else (TUMTRAF_CLASSES if coop else NUSC_CLASSES))
    if final_kernel is None:
        final_kernel = 3 if 'Lidar' in kind else 1
    post = [-80, -80, -10.0, 80, 80, 10.0] if coop else [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
    return dict(
        type=kind,
        in_channels=in_channels,
        hidden_dim=hidden_dim,
        num_query=num_query,
        downsample_scale=8,
        common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2)),
        tasks=[dict(num_class=len(classes), class_names=classes)],
        bbox_coder=dict(type='MultiTaskBBoxCoder', post_center_range=post, pc_range=pc_range, max_num=max_num,
                        voxel_size=[0.075, 0.075, 0.2], num_classes=len(classes)),
        separate_head=dict(type='SeparateTaskHead', init_bias=-2.19, final_kernel=final_kernel),
        transformer=dict(
            type=_TRANSFORMER_OF[kind],
            decoder=dict(
                type='PETRTransformerDecoder', return_intermediate=True, num_layers=num_layers,
                transformerlayers=dict(
                    type='PETRTransformerDecoderLayer', with_cp=False,
                    attn_cfgs=[
                        dict(type='MultiheadAttention', embed_dims=hidden_dim, num_heads=8, dropout=0.1),
                        dict(type=cross_attn, embed_dims=hidden_dim, num_heads=8, dropout=0.1),
                    ],
                    ffn_cfgs=dict(type='FFN', embed_dims=hidden_dim, feedforward_channels=4 * hidden_dim, num_fcs=2,
                                  ffn_drop=0., act_cfg=dict(type='ReLU', inplace=True)),
                    feedforward_channels=4 * hidden_dim,
                    operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2, alpha=0.25, reduction='mean', loss_weight=2.0),
        loss_bbox=dict(type='L1Loss', reduction='mean', loss_weight=0.25),
        loss_heatmap=dict(type='GaussianFocalLoss', reduction='mean', loss_weight=1.0),
        train_cfg=None,
        test_cfg=dict(grid_size=[grid, grid, 40]),
    )


# ------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------
def _rng(key: str, seed: int):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synth_tensor(key: str, shape, seed: int = 0) -> np.ndarray:
    """Reference-style initial value for state-dict entry `key` (float32 unless noted)."""
    shape = tuple(int(s) for s in shape)
    r = _rng(key, seed)
    leaf = key.split('.')[-1]
    if leaf == 'num_batches_tracked':
        return np.asarray(100, dtype=np.int64)
    if leaf == 'running_mean':
        return r.normal(0, 0.1, shape).astype(np.float32)
    if leaf == 'running_var':
        return r.uniform(0.5, 1.5, shape).astype(np.float32)
    if 'reference_points' in key:
        return r.uniform(0, 1, shape).astype(np.float32)           # cmt_head.py:320-322
    is_norm = ('.norms.' in key or 'post_norm' in key or '.bn.' in key or
               (len(shape) == 1 and leaf == 'weight'))             # LN / BN / GroupLayerNorm1d affine
    if is_norm and leaf == 'weight':
        return (1.0 + r.normal(0, 0.05, shape)).astype(np.float32)
    if leaf in ('bias', 'in_proj_bias'):
        if 'cls_logits' in key and key.split('.')[-2] == '3':
            return np.full(shape, -2.19, dtype=np.float32) + r.normal(0, 0.02, shape).astype(np.float32)  # :164-172
        return r.normal(0, 0.02, shape).astype(np.float32)
    if 'task_heads' in key:                                        # Kaiming normal, fan_out (mmcv 'Kaiming')
        fan_out = shape[0] * (shape[2] if len(shape) > 2 else 1)
        return r.normal(0, np.sqrt(2.0 / fan_out), shape).astype(np.float32)
    if 'transformer' in key:                                       # xavier uniform (cmt_transformer.py:77-82)
        fan_out, fan_in = shape[0], int(np.prod(shape[1:]))
        if leaf == 'in_proj_weight':
            fan_out = shape[0]                                     # attention.py:120-124 xavier on [3E, E]
        a = np.sqrt(6.0 / (fan_in + fan_out))
        return r.uniform(-a, a, shape).astype(np.float32)
    if len(shape) == 4:                                            # shared_conv: kaiming-ish
        fan_in = int(np.prod(shape[1:]))
        return r.normal(0, np.sqrt(2.0 / fan_in), shape).astype(np.float32)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]  # nn.Linear default: U(-1/sqrt(fan_in), ..)
    a = 1.0 / np.sqrt(fan_in)
    return r.uniform(-a, a, shape).astype(np.float32)


def synth_state_dict(shapes: dict, seed: int = 0) -> dict:
    """shapes: {state-dict key: shape}. Returns {key: np.ndarray}."""
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def load_synth_weights(module, seed: int = 0):
    """Fill a torch module (reference or ours -- they share state-dict keys) in place."""
    import torch
    sd = module.state_dict()
    new = {k: torch.from_numpy(np.asarray(synth_tensor(k, v.shape, seed))).to(v.dtype).reshape(v.shape)
           for k, v in sd.items()}
    module.load_state_dict(new, strict=True)
    return module


# ------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------
def camera_matrices(n_views: int, rng, pad_w: float, pad_h: float, t_sigma: float = 0.3):
    """lidar2img = K @ E_v: pinhole intrinsics scaled to the padded image, E_v = camera-axes swap
    x yaw 2*pi*v/V + small translation (invertible, well conditioned)."""
    fx = 1266.0 * pad_w / 1600.0
    K = np.array([[fx, 0, pad_w / 2, 0], [0, fx, pad_h / 2, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
    swap = np.array([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1]], dtype=np.float64)  # lidar xyz -> cam
    mats = []
    for v in range(n_views):
        yaw = 2 * np.pi * v / max(n_views, 1)
        c, s = np.cos(yaw), np.sin(yaw)
        R = np.array([[c, s, 0, 0], [-s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
        T = np.eye(4)
        T[:3, 3] = rng.normal(0, t_sigma, 3)
        mats.append(K @ swap @ R @ T)
    return mats


def rigid_transform(rng):
    yaw = rng.uniform(-np.pi, np.pi)
    c, s = np.cos(yaw), np.sin(yaw)
    M = np.eye(4)
    M[:2, :2] = [[c, -s], [s, c]]
    M[:3, 3] = rng.normal(0, 5.0, 3) * [1, 1, 0.1]
    return M


def make_inputs(kind='CmtHead', *, B=1, bev_hw=180, n_views=6, img_hw=(40, 100), in_channels=512, hidden_dim=256,
                vehicle_views=1, infra_views=3, seed=0, stride=16):
    """Synthetic head inputs as numpy arrays + img_metas.
    Returns dict(pts_feats=..., img_feats=..., img_metas=[...]) -- for coop heads the four feature
    entries are vehicle_pts_feats / infrastructure_pts_feats / vehicle_img_feats / infrastructure_img_feats."""
    rng = np.random.RandomState(1000 + seed)
    h, w = img_hw
    pad_h, pad_w = h * stride, w * stride
    has_pts = 'Image' not in kind
    has_img = 'Lidar' not in kind
    coop = kind.endswith('Coop')
    metas = [dict(box_type_3d=None) for _ in range(B)]
    out = dict(img_metas=metas, pad_shape=(pad_h, pad_w, 3))

    def feats(n, c, hh, ww):
        return rng.standard_normal((n, c, hh, ww)).astype(np.float32)

    if not coop:
        out['pts_feats'] = feats(B, in_channels, bev_hw, bev_hw) if has_pts else None
        out['img_feats'] = feats(B * n_views, hidden_dim, h, w) if has_img else None
        if has_img:
            for m in metas:
                m['lidar2img'] = camera_matrices(n_views, rng, pad_w, pad_h)
                m['pad_shape'] = [(pad_h, pad_w, 3)] * n_views
    else:
        for node, nv in (('vehicle', vehicle_views), ('infrastructure', infra_views)):
            out[f'{node}_pts_feats'] = feats(B, in_channels, bev_hw, bev_hw) if has_pts else None
            out[f'{node}_img_feats'] = feats(B * nv, hidden_dim, h, w) if has_img else None
            if has_img:
                for m in metas:
                    mats = camera_matrices(nv, rng, pad_w, pad_h)
                    if node == 'vehicle':  # transforms_3d_coop.py:213-222: lidar2img @ inv(v2i)
                        v2i = rigid_transform(rng)
                        mats = [M @ np.linalg.inv(v2i) for M in mats]
                    m[f'{node}_lidar2img'] = mats
                    m[f'{node}_pad_shape'] = [(pad_h, pad_w, 3)] * nv
    return out


def mini_case(kind: str, seed: int = 0):
    """Small but architecturally exact case (C=256, 8 heads, FFN 1024, 64 depth bins) used for the
    committed golden vectors: 2 decoder layers, 96 queries, 24x24 BEV, 6x10 image features."""
    cfg = head_cfg(kind, num_query=96, num_layers=2, grid=8 * 24, max_num=50)
    inputs = make_inputs(kind, B=2, bev_hw=24, n_views=2, img_hw=(6, 10), vehicle_views=1, infra_views=2, seed=seed)
    return cfg, inputs


def deep_copy_cfg(cfg):
    return copy.deepcopy(cfg)
