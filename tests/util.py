"""Shared helpers for the test-suite."""
import json
import os

import numpy as np
import torch

from ciaosr_b200 import synth
from ciaosr_b200.builder import build
from ciaosr_b200.generators import LocalImplicitSREDSR

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

HEAD_CASES = ["head_small", "head_frac", "head_ls1", "head_ls3", "head_nonl0", "head_c64",
              "head_c64_nonl0", "head_c180", "head_c180_nonl0"]
CSATTN_CASES = ["csattn_c64", "csattn_odd", "csattn_c180"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return meta, arrays


def generator_cfg(c, hidden, eval_bsize=None, local_size=2, non_local=True, num_blocks=1, **extra):
    mlp = lambda: dict(type="MLPRefiner", in_dim=4, out_dim=3, hidden_list=list(hidden))
    return dict(type=LocalImplicitSREDSR,
                encoder=dict(type="EDSR", in_channels=3, out_channels=3, mid_channels=c,
                             num_blocks=num_blocks),
                imnet_q=mlp(), imnet_k=mlp(), imnet_v=mlp(), local_size=local_size,
                feat_unfold=True, eval_bsize=eval_bsize, non_local_attn=non_local, **extra)


def build_generator(meta, device="cpu", **extra):
    """This package's generator with the same synthetic weights the golden run used."""
    g = build(generator_cfg(meta["c"], meta["hidden"], meta.get("eval_bsize"),
                            meta.get("local_size", 2), meta.get("non_local", True), **extra))
    synth.fill_module(g, meta["seed"])
    return g.eval().to(device)


def head_weights(gen):
    return {k: v.detach().cpu() for k, v in gen.state_dict().items()
            if k.startswith(("imnet_", "cs_attn."))}


def max_abs(a, b):
    return float((a.double() - b.double()).abs().max())
