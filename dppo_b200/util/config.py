"""Attribute-style nested config (the subset of OmegaConf's DictConfig the agent loop touches: cfg.a.b, cfg.get)."""

import importlib


class Cfg(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def instantiate(node, **overrides):
    """hydra.utils.instantiate for a {_target_: "pkg.mod.Class", **kwargs} node (nested nodes built first)."""
    kw = {}
    for k, v in node.items():
        if k == "_target_":
            continue
        kw[k] = instantiate(v) if isinstance(v, dict) and "_target_" in v else v
    kw.update(overrides)
    mod, _, name = node["_target_"].rpartition(".")
    return getattr(importlib.import_module(mod), name)(**kw)
