"""Registries behind which the reference selects its classes by name (`type='CmtHead'`, ...).

The reference registers into mmcv / mmdet registries (e.g. `@HEADS.register_module()`,
models/dense_heads/cmt_head.py:206; `@TRANSFORMER.register_module()`, models/utils/cmt_transformer.py:48;
`@ATTENTION.register_module()`, models/utils/petr_transformer.py:37,182).  Those packages are not
installed here, so the same decorator API is provided by a small built-in registry; when mmcv/mmdet
*are* importable, `register_into_openmmlab()` mirrors every class into the real registries so that
`plugin_dir` loading of the unmodified configs picks up these classes.
"""
from __future__ import annotations

import copy


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self._modules and not force and self._modules[key] is not cls:
                raise KeyError(f"{key} already registered in {self.name}")
            self._modules[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        if key not in self._modules:
            raise KeyError(f"{key!r} is not registered in the {self.name} registry "
                           f"(known: {sorted(self._modules)})")
        return self._modules[key]

    def build(self, cfg, **default_args):
        cfg = dict(copy.deepcopy(cfg))
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        typ = cfg.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        return cls(**cfg)

    def __contains__(self, key):
        return key in self._modules

    def keys(self):
        return self._modules.keys()


HEADS = Registry("HEADS")
TRANSFORMER = Registry("TRANSFORMER")
ATTENTION = Registry("ATTENTION")
FEEDFORWARD_NETWORK = Registry("FEEDFORWARD_NETWORK")
TRANSFORMER_LAYER = Registry("TRANSFORMER_LAYER")
TRANSFORMER_LAYER_SEQUENCE = Registry("TRANSFORMER_LAYER_SEQUENCE")
BBOX_CODERS = Registry("BBOX_CODERS")

ALL = dict(HEADS=HEADS, TRANSFORMER=TRANSFORMER, ATTENTION=ATTENTION,
           FEEDFORWARD_NETWORK=FEEDFORWARD_NETWORK, TRANSFORMER_LAYER=TRANSFORMER_LAYER,
           TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE, BBOX_CODERS=BBOX_CODERS)


class ConfigDict(dict):
    """Attribute-access dict (the head reads `transformer.decoder.num_layers`, cmt_head.py:310)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v


def register_into_openmmlab(force=True):
    """Mirror every class into mmcv/mmdet registries when the OpenMMLab stack is importable.
    Returns the list of (registry, name) pairs registered; [] when mmcv/mmdet are absent."""
    done = []
    try:
        from mmcv.cnn.bricks.registry import (ATTENTION as A, TRANSFORMER_LAYER as TL,
                                              TRANSFORMER_LAYER_SEQUENCE as TLS)
        from mmdet.core.bbox.builder import BBOX_CODERS as BC
        from mmdet.models import HEADS as H
        from mmdet.models.utils.builder import TRANSFORMER as T
    except Exception:
        return done
    targets = dict(HEADS=H, TRANSFORMER=T, ATTENTION=A, TRANSFORMER_LAYER=TL,
                   TRANSFORMER_LAYER_SEQUENCE=TLS, BBOX_CODERS=BC)
    for rname, target in targets.items():
        for key, cls in ALL[rname]._modules.items():
            if key in ("MultiheadAttention", "FFN"):
                continue  # mmcv's own classes stay in place
            target.register_module(name=key, force=force, module=cls)
            done.append((rname, key))
    return done
