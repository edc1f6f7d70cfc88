"""Diagnostic: error of both engines vs the CPU oracle for the SwinIR-sized head (C = 180) as the feature
magnitude grows (the head amplifies rounding through the inner-attention logits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ciaosr_b200 import synth
from ciaosr_b200.coords import make_cell, make_coord
from oracle import ciaosr_oracle as orc
from tests.util import build_generator, head_weights, max_abs

dev = torch.device("cuda:0")
for c, non_local in [(180, False), (180, True), (64, True)]:
    meta = dict(c=c, hidden=[256] * 4, eval_bsize=None, local_size=2, non_local=non_local, seed=3)
    g = build_generator(meta, dev)
    w = head_weights(g)
    for sigma in (0.5, 0.8, 1.2):
        b, h, wd, s = 1, 12, 12, 4
        feat = synth.synth_feature(b, c, h, wd, 5) * (sigma / 0.5)
        lq = synth.synth_lr_image(b, h, wd, 5)
        coord = make_coord((h * s, wd * s)).unsqueeze(0)
        cell = make_cell((h * s, wd * s), coord.shape[1]).unsqueeze(0)
        ref = orc.head_forward(lq, feat, coord, cell, w, eval_bsize=None, non_local_attn=non_local)
        plan = g.head_plan()
        res = {}
        for eng in ("simt", "tcgen05"):
            out = plan.query_rgb(feat.to(dev), coord.to(dev), cell.to(dev), lr_image=lq.to(dev), engine=eng).cpu()
            res[eng] = max_abs(out, ref)
        print(f"C={c} non_local={non_local} sigma={sigma}: |ref|max {float(ref.abs().max()):.2f} "
              f"simt-oracle {res['simt']:.2e} tcgen05-oracle {res['tcgen05']:.2e}", flush=True)
