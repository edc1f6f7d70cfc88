#!/usr/bin/env python
"""Non-gating error-vs-feature-magnitude table: max-abs of both engines against the CPU oracle
(fp32, pinned to the reference) as the encoder features are scaled to std sigma = 0.25 ... 2, at the real head
dimensions (C = 64, hidden 256x4, cross-scale attention on, 24x20 LR -> x4).  For each sigma also: the error of the
cross-scale attention alone, of the head given the oracle's non-local map (so the stage that loses accuracy is
visible), the oracle's own float64-vs-float32 distance (the noise floor of "the reference in fp32"), and the output
range.  Writes gpurun_out/sigma_table.md.

    python tools/sigma_table.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ciaosr_b200 import synth  # noqa: E402
from ciaosr_b200.coords import make_cell, make_coord  # noqa: E402
from oracle import ciaosr_oracle as orc  # noqa: E402
from tests.util import build_generator, head_weights, max_abs  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    meta = dict(c=64, hidden=[256, 256, 256, 256], eval_bsize=30000, local_size=2, non_local=True, seed=77)
    g = build_generator(meta, dev)
    w = head_weights(g)
    w64 = {k: v.double() for k, v in w.items()}
    b, h, wd, s = 1, 24, 20, 4
    base = synth.synth_feature(b, 64, h, wd, 77) / 0.5            # unit std
    lq = synth.synth_lr_image(b, h, wd, 77)
    coord = make_coord((h * s, wd * s)).unsqueeze(0)
    cell = make_cell((h * s, wd * s), coord.shape[1]).unsqueeze(0)
    plan = g.head_plan()
    rows = []
    for sigma in (0.25, 0.4, 0.5, 0.75, 1.0, 1.5, 2.0):
        feat = (base * sigma).contiguous()
        ref = orc.head_forward(lq, feat, coord, cell, w, eval_bsize=30000)
        nl_ref = orc.cross_scale_attention(feat, w)
        try:
            orc.F32 = torch.float64
            ref64 = orc.head_forward(lq.double(), feat.double(), coord.double(), cell.double(), w64, eval_bsize=30000)
            floor = max_abs(ref, ref64)
            if floor > 1e-2:                                   # a gather index rounded differently in float64
                floor = float("nan")
        except Exception:                                       # the oracle is fp32-typed in places
            floor = float("nan")
        finally:
            orc.F32 = torch.float32
        rec = dict(sigma=sigma, out_absmax=float(ref.abs().max()), floor=floor)
        fd, cd, ld, qd = feat.to(dev), coord.to(dev), cell.to(dev), lq.to(dev)
        for engine in ("tcgen05", "simt"):
            nl = plan.cross_scale_attention(fd, engine=engine)
            out = plan.query_rgb(fd, cd, ld, lr_image=qd, nonlocal_feat=nl, eval_bsize=30000, engine=engine)
            out_h = plan.query_rgb(fd, cd, ld, lr_image=qd, nonlocal_feat=nl_ref.to(dev), eval_bsize=30000, engine=engine)
            rec[engine] = max_abs(out.cpu(), ref)
            rec[engine + "_csattn"] = max_abs(nl.cpu(), nl_ref)
            rec[engine + "_head_given_nl"] = max_abs(out_h.cpu(), ref)
        rows.append(rec)
        print(rec, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sigma_table.md"), "w") as f:
        f.write("| feature std | max abs(out) | oracle f32 vs f64 | tcgen05 vs oracle | - cs-attn alone | - head given oracle nl "
                "| fp32 SIMT vs oracle | - cs-attn alone | - head given oracle nl |\n|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            f.write(f"| {r['sigma']} | {r['out_absmax']:.2f} | {r['floor']:.1e} | {r['tcgen05']:.1e} | {r['tcgen05_csattn']:.1e} | "
                    f"{r['tcgen05_head_given_nl']:.1e} | {r['simt']:.1e} | {r['simt_csattn']:.1e} | {r['simt_head_given_nl']:.1e} |\n")


if __name__ == "__main__":
    main()
