"""Mint the BASELINE-size golden vectors by RUNNING THE REFERENCE (build container only).

    python -m oracle.make_golden_full [case ...]      # rewrites tests/golden/full_*.npz

`oracle/make_golden.py` pins the path on small shapes; this script pins it at the sizes
BASELINE.json's configs are quoted on, where the tcgen05 engine's long-K accumulation and
the full-depth encoder matter ("no BASELINE-size case is ever compared with
the reference").  Every case runs the UNMODIFIED reference modules imported in place from
/root/reference (oracle/ref_harness.py) on the CPU in fp32:

  full_cfg2        config 2 end to end: LocalImplicitSRRDN (RDN 16 blocks x 8 layers, growth 64,
                   imnet 256x4, cross-scale attention), crops 0 and 1 of bench.py's batch of 16
                   48x48 LR crops -> x4, `forward(test_mode=True)` with eval_bsize 30000
                   (ciaosr_net.py:88-110, 226-248).  Also stores the encoder's feature map of crop 0.
  full_csattn_c64  CrossScaleAttention.forward on one 192x192 map at C = 64 (arch_csnln.py:430-532);
                   a strided subset of the output is stored (every 5th pixel in y and x).
  full_csattn_c180 the same at C = 180 on 96x96.
  full_cfg3_x3     one config-3 tile: RDN 16x8 on a 192x192 LR tile -> x3 (331 776 queries); the
                   reference recomputes the 192x192 cross-scale attention per eval_bsize chunk, so the
                   chunk length is raised to 120 000 (3 chunks; the result does not depend on it because
                   every cell is equal).  Every 7th query is stored.
  full_frac        fractional scales x2.5 / x1.7 with the real head dimensions (C = 64, hidden 256x4), so
                   that the tcgen05 engine is pinned on non-integer scales too.

Inputs are regenerated from seeds (ciaosr_b200/synth.py, bench.make_inputs) by the tests; only the
reference's outputs (and index lists) are stored.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ciaosr_b200 import synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle.make_golden import grid_coord_cell, save  # noqa: E402

RGB_MEAN = (0.4488, 0.4371, 0.4040)
HIDDEN = (256, 256, 256, 256)


def rdn_reference(eval_bsize):
    g = rh.build_reference_generator("rdn", 64, HIDDEN, num_blocks=16, num_layers=8, eval_bsize=eval_bsize)
    synth.fill_module(g, 0)
    return g


def cfg2():
    import bench
    g = rdn_reference(30000)
    lq, coord, cell = bench.make_inputs(16, 100)           # rank 0's batch
    lq = (lq - torch.tensor(RGB_MEAN).view(1, 3, 1, 1))[:2].contiguous()
    coord, cell = coord[:2].contiguous(), cell[:2].contiguous()
    t0 = time.time()
    with torch.no_grad():
        feat = g.gen_feature(lq)[0]
        out = g(lq, coord, cell, test_mode=True)
    print(f"cfg2: {time.time() - t0:.1f} s, feature std {float(feat.std()):.3f} max {float(feat.abs().max()):.2f}, "
          f"out range [{float(out.min()):.2f}, {float(out.max()):.2f}]")
    save("full_cfg2", dict(kind="full_cfg2", crops=2, seed=100, feature_std=float(feat.std()),
                           feature_absmax=float(feat.abs().max())),
         dict(out=out.numpy(), feature0=feat[0].numpy()))


def csattn(name, c, n, stride, seed):
    ref = rh.import_reference()
    m = ref.csnln.CrossScaleAttention(channel=c, scale=[2]).eval()
    holder = torch.nn.Module()
    holder.cs_attn = m
    synth.fill_module(holder, seed)
    feature = synth.synth_feature(1, c, n, n, seed)
    t0 = time.time()
    with torch.no_grad():
        out = m(feature)
    print(f"{name}: {time.time() - t0:.1f} s, out std {float(out.std()):.3f}")
    sub = out[:, :, ::stride, ::stride].contiguous()
    save(name, dict(kind="full_csattn", c=c, n=n, stride=stride, seed=seed, out_std=float(out.std())),
         dict(out_sub=sub.numpy()))


def cfg3_x3():
    g = rdn_reference(120000)
    n, s, every = 192, 3, 7
    lq = synth.synth_lr_image(1, n, n, 7)
    coord, cell = grid_coord_cell(1, n, n, s)
    t0 = time.time()
    with torch.no_grad():
        out = g(lq, coord, cell, test_mode=True)
    print(f"cfg3_x3: {time.time() - t0:.1f} s, out range [{float(out.min()):.2f}, {float(out.max()):.2f}]")
    save("full_cfg3_x3", dict(kind="full_cfg3", n=n, scale=s, every=every, seed=7),
         dict(out_sub=out[:, ::every].contiguous().numpy()))


def frac():
    g = rh.build_reference_generator("edsr", 64, HIDDEN, num_blocks=1, eval_bsize=900)
    synth.fill_module(g, 31)
    b, h, w = 2, 11, 13
    feature = synth.synth_feature(b, 64, h, w, 31)
    x_lr = synth.synth_lr_image(b, h, w, 31)
    g.gen_feature = lambda _x: [feature]
    arrays, tags = {}, []
    with torch.no_grad():
        for s in (2.5, 1.7):
            coord, cell = grid_coord_cell(b, h, w, s)
            tag = f"s{s}"
            arrays[f"coord_{tag}"], arrays[f"cell_{tag}"] = coord.numpy(), cell.numpy()
            arrays[f"out_{tag}"] = g(x_lr, coord, cell, test_mode=True).numpy()
            tags.append(tag)
    save("full_frac", dict(kind="full_frac", c=64, hidden=list(HIDDEN), b=b, h=h, w=w, seed=31, eval_bsize=900,
                           local_size=2, non_local=True, tags=tags), arrays)


CASES = {
    "cfg2": cfg2,
    "csattn_c64": lambda: csattn("full_csattn_c64", 64, 192, 5, 41),
    "csattn_c180": lambda: csattn("full_csattn_c180", 180, 96, 3, 42),
    "cfg3_x3": cfg3_x3,
    "frac": frac,
}


def main(argv):
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for name in (argv or list(CASES)):
        CASES[name]()


if __name__ == "__main__":
    main(sys.argv[1:])
