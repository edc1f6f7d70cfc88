"""Deterministic synthetic weights / inputs (no checkpoints or datasets ship
with the reference: pretrain_model/put_models_here and data/put_data_here are
empty files).

Weights are drawn from ``numpy.random.RandomState`` (bit-stable across numpy
versions and machines) keyed on the state_dict key, so any two modules that
expose the same keys and shapes -- the reference generator in the build
container and this package's generator on the GPU box -- get identical
parameters without a checkpoint file travelling between them.

The scales are chosen so that the head is *numerically live*: hidden
activations stay O(1) through the ReLU stacks and the four inner-attention
logits differ by O(1), so the softmax over neighbours is far from uniform.
(With PyTorch's default init the logits are ~1e-3 and a wrong attention
weight would hide under a 1e-4 tolerance.)
"""
import zlib

import numpy as np
import torch


def _rs(seed, key):
    return np.random.RandomState((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))


def synth_tensor(key, shape, seed=0, gain=1.4):
    """Value for one state_dict entry (see module docstring)."""
    rs = _rs(seed, key)
    shape = tuple(shape)
    if key.endswith("escape_NaN"):
        return torch.full(shape, 1e-4, dtype=torch.float32)
    if key.endswith("relative_position_index") or key.endswith("attn_mask"):
        return None                                   # integer / derived buffers: keep
    if key.endswith(".bias"):
        return torch.from_numpy(rs.normal(0.0, 0.05, size=shape).astype(np.float32))
    if len(shape) == 1:
        if "norm" in key:                              # LayerNorm gain
            return torch.from_numpy((1.0 + rs.normal(0.0, 0.02, size=shape)).astype(np.float32))
        # PReLU slope (cs_attn.conv_*.1.weight has shape [1])
        return torch.from_numpy((0.25 + rs.uniform(-0.1, 0.1, size=shape)).astype(np.float32))
    fan_in = int(np.prod(shape[1:]))
    std = gain / np.sqrt(fan_in)
    return torch.from_numpy(rs.normal(0.0, std, size=shape).astype(np.float32))


def fill_module(module, seed=0):
    """Overwrite every floating parameter/buffer of `module` in place."""
    sd = module.state_dict()
    # the final Linear of imnet_k gets a smaller gain: keeps the inner-attention
    # logits (a 9C-term sum of query*key*weight) O(1) apart instead of O(10)
    last = {}
    for k in sd:
        if ".layers." in k and k.endswith(".weight"):
            pre, idx = k.split(".layers.")
            i = int(idx.split(".")[0])
            last[pre] = max(last.get(pre, -1), i)
    small = {f"{pre}.layers.{i}.weight" for pre, i in last.items() if pre.endswith("imnet_k")}
    with torch.no_grad():
        for k, v in sd.items():
            if not torch.is_floating_point(v):
                continue
            if k in small:
                gain = 0.35
            elif k.startswith(("imnet_", "cs_attn.")) or ".imnet_" in k or ".cs_attn." in k:
                gain = 1.4
            else:
                gain = 0.8        # encoder convolutions: keeps the 130-conv RDN's features O(1)
            t = synth_tensor(k, v.shape, seed, gain=gain)
            if t is not None:
                v.copy_(t.to(v.dtype))
    return module


def synth_feature(b, c, h, w, seed=0):
    rs = _rs(seed, f"feature{b}x{c}x{h}x{w}")
    return torch.from_numpy(rs.normal(0.0, 0.5, size=(b, c, h, w)).astype(np.float32))


def synth_lr_image(b, h, w, seed=0, mean=(0.4488, 0.4371, 0.4040)):
    """U[0,1) LR image, normalised like CiaoSR.forward_test (ciaosr.py:142-144, std=1)."""
    rs = _rs(seed, f"lq{b}x{h}x{w}")
    img = torch.from_numpy(rs.uniform(0.0, 1.0, size=(b, 3, h, w)).astype(np.float32))
    return img - torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
