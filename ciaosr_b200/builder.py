"""mmcv/mmedit-free builders for the reference's config dicts.

The reference's configs pass class objects as ``type`` for its own classes and
strings for upstream mmedit ones (configs/001_...rdn...py:13-45: 'RDN', 'EDSR',
'MLPRefiner', 'L1Loss').  mmcv / mmedit are not installable here, so the string
names resolve in this small registry instead.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

_REGISTRY = {}


def register(name, obj=None):
    if obj is None:
        return lambda o: register(name, o)
    _REGISTRY[name] = obj
    return obj


def _resolve(typ):
    if isinstance(typ, str):
        if not _REGISTRY:
            _populate()
        if typ not in _REGISTRY:
            raise KeyError(f"{typ} is not in the ciaosr_b200 registry ({sorted(_REGISTRY)})")
        return _REGISTRY[typ]
    return typ


def build(cfg, **default_args):
    """Instantiate ``cfg['type']`` (class object or registered name) with the other keys."""
    if not isinstance(cfg, dict) or "type" not in cfg:
        raise TypeError(f"cfg must be a dict with a 'type' key, got {type(cfg)}")
    args = {k: v for k, v in cfg.items() if k != "type"}
    for k, v in default_args.items():
        args.setdefault(k, v)
    return _resolve(cfg["type"])(**args)


build_backbone = build
build_component = build
build_model = build


class L1Loss(nn.Module):
    """mmedit's L1Loss (mean/sum/none reduction, loss_weight)."""

    def __init__(self, loss_weight=1.0, reduction="mean", sample_wise=False):
        super().__init__()
        if reduction not in ("none", "mean", "sum"):
            raise ValueError(f"Unsupported reduction mode: {reduction}")
        self.loss_weight, self.reduction = loss_weight, reduction

    def forward(self, pred, target, weight=None, **kwargs):
        loss = (pred - target).abs()
        if weight is not None:
            loss = loss * weight
        if self.reduction == "mean":
            loss = loss.mean()
        elif self.reduction == "sum":
            loss = loss.sum()
        return self.loss_weight * loss


def build_loss(cfg):
    return build(cfg)


def _populate():
    from . import encoders, pipelines, refiners, swinir
    _REGISTRY.update({"RDN": encoders.RDN, "EDSR": encoders.EDSR, "SwinIR": swinir.SwinIR,
                      "MLPRefiner": refiners.MLPRefiner, "L1Loss": L1Loss,
                      "GenerateCoordinateAndCell": pipelines.GenerateCoordinateAndCell})


def load_checkpoint(module, filename, map_location="cpu", strict=False, revise_keys=((r"^module\.", ""),)):
    """Minimal stand-in for mmcv.runner.load_checkpoint (tools/test.py:115-118)."""
    import re
    ckpt = torch.load(filename, map_location=map_location)
    state = ckpt.get("state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    for pat, rep in revise_keys:
        state = {re.sub(pat, rep, k): v for k, v in state.items()}
    missing, unexpected = module.load_state_dict(state, strict=strict)
    return dict(missing_keys=missing, unexpected_keys=unexpected)


class Config(dict):
    """``mmcv.Config.fromfile`` for python config files: executes the file and keeps its
    public module-level names (tools/test.py:72)."""

    def __getattr__(self, name):
        try:
            v = self[name]
        except KeyError as e:
            raise AttributeError(name) from e
        return v

    @classmethod
    def fromfile(cls, filename):
        filename = os.path.abspath(filename)
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        name = "_ciaosr_cfg_" + os.path.splitext(os.path.basename(filename))[0].replace(".", "_")
        spec = importlib.util.spec_from_file_location(name, filename)
        mod = importlib.util.module_from_spec(spec)
        # configs import `mmedited.*`; make sure this repo's compat package is importable
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if root not in sys.path:
            sys.path.insert(0, root)
        spec.loader.exec_module(mod)
        cfg = cls()
        for k, v in vars(mod).items():
            if k.startswith("_") or isinstance(v, types.ModuleType) or isinstance(v, type) \
                    or callable(v):
                continue
            cfg[k] = v
        cfg["filename"] = filename
        return cfg
