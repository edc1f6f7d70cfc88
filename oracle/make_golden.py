"""Mint golden vectors by RUNNING THE REFERENCE (build container only).

    python -m oracle.make_golden            # rewrites tests/golden/*.npz

The reference ships no tests and no golden vectors, so
parity is pinned on outputs of the reference's own code, imported in place from
/root/reference through oracle/ref_harness.py, on seeded synthetic weights and
inputs (ciaosr_b200/synth.py; bit-stable numpy RandomState streams keyed on the
state_dict key, so the weights themselves need not be stored).

Each .npz holds the case's `meta` (JSON), its inputs and the reference's outputs.
`/root/reference` does not travel to the GPU box; these files do.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ciaosr_b200 import synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def grid_coord_cell(b, h, w, s):
    th, tw = round(h * s), round(w * s)
    coord = rh.make_coord((th, tw)).unsqueeze(0).expand(b, -1, 2).contiguous()
    cell = torch.ones_like(coord)
    cell[:, :, 0] *= 2 / th
    cell[:, :, 1] *= 2 / tw
    return coord, cell


def head_case(name, c, hidden, b, h, w, scales, eval_bsize, local_size=2, non_local=True,
              seed=0, random_q=0):
    g = rh.build_reference_generator("edsr", c, hidden, num_blocks=1, eval_bsize=eval_bsize,
                                     local_size=local_size, non_local_attn=non_local)
    synth.fill_module(g, seed)
    feature = synth.synth_feature(b, c, h, w, seed)
    x_lr = synth.synth_lr_image(b, h, w, seed)
    g.gen_feature = lambda _x: [feature]
    arrays = dict(feature=feature.numpy(), x_lr=x_lr.numpy())
    with torch.no_grad():
        if non_local:
            arrays["nonlocal"] = g.cs_attn(feature).numpy()
        runs = []
        for s in scales:
            coord, cell = grid_coord_cell(b, h, w, s)
            runs.append((f"s{s}", coord, cell))
        if random_q:
            rs = np.random.RandomState(seed + 17)
            coord = torch.from_numpy(rs.uniform(-1, 1, size=(b, random_q, 2)).astype(np.float32))
            sc = rs.uniform(1.0, 4.0, size=(b, 1, 1)).astype(np.float32)
            cell = torch.from_numpy(np.broadcast_to(
                np.concatenate([2.0 / (h * sc), 2.0 / (w * sc)], axis=2), (b, random_q, 2)).copy())
            runs.append(("rand", coord, cell))
        for tag, coord, cell in runs:
            arrays[f"coord_{tag}"] = coord.numpy()
            arrays[f"cell_{tag}"] = cell.numpy()
            arrays[f"pred_{tag}"] = g.query_rgb([feature], coord, cell).numpy()        # bare head, one chunk
            arrays[f"out_{tag}"] = g(x_lr, coord, cell, test_mode=True).numpy()         # chunked + residual
    meta = dict(kind="head", c=c, hidden=list(hidden), b=b, h=h, w=w, eval_bsize=eval_bsize,
                local_size=local_size, non_local=non_local, seed=seed,
                tags=[t for t, _, _ in runs])
    save(name, meta, arrays)


def csattn_case(name, c, b, h, w, seed=0):
    ref = rh.import_reference()
    m = ref.csnln.CrossScaleAttention(channel=c, scale=[2]).eval()
    holder = torch.nn.Module()
    holder.cs_attn = m
    synth.fill_module(holder, seed)
    feature = synth.synth_feature(b, c, h, w, seed)
    with torch.no_grad():
        out = m(feature)
    save(name, dict(kind="csattn", c=c, b=b, h=h, w=w, seed=seed),
         dict(feature=feature.numpy(), out=out.numpy()))


def clip_case(name, c, hidden, h, w, scale, tile, overlap, seed=0):
    ref = rh.import_reference()
    g = rh.build_reference_generator("edsr", c, hidden, num_blocks=1, eval_bsize=500)
    synth.fill_module(g, seed)
    lq = synth.synth_lr_image(1, h, w, seed)
    fake_self = types.SimpleNamespace(test_cfg=dict(scale=scale, tile=tile, tile_overlap=overlap))
    with torch.no_grad():
        out = ref.restorer.CiaoSR.clip_test(fake_self, lq, g)
    save(name, dict(kind="clip", c=c, hidden=list(hidden), h=h, w=w, scale=scale, tile=tile,
                    overlap=overlap, seed=seed, eval_bsize=500),
         dict(lq=lq.numpy(), out=out.numpy()))


def real_clip_case(name, c, hidden, h, w, scale, tile, overlap, seed=0):
    """`RealCiaoSR.clip_test` (real_ciaosr.py:336-373) on a head without cross-scale attention.  The method is taken from the reference file by name (the module itself needs mmedit's GAN
    classes at import time) and run unbound; its two `.cuda()` calls are made no-ops for the CPU run."""
    import ast
    src = "/root/reference/mmedited/models/restorers/real_ciaosr.py"
    tree = ast.parse(open(src).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RealCiaoSR")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "clip_test")
    ns = {"torch": torch, "make_coord": rh.make_coord}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src, "exec"), ns)
    # NB the reference's generator classes do not accept the extra keywords its own 002 configs pass
    # (local_ensemble_coord, imnet_k_type, imnet_v_type, res, cat_nla_v -> TypeError at this commit), so the only
    # 002 head flag that can be exercised against the reference is non_local_attn=False; the residual stays on.
    g = rh.build_reference_generator("edsr", c, hidden, num_blocks=1, eval_bsize=500, non_local_attn=False)
    synth.fill_module(g, seed)
    lq = synth.synth_lr_image(1, h, w, seed)
    fake_self = types.SimpleNamespace(test_cfg=dict(scale=scale, tile=tile, tile_overlap=overlap))
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            out = ns["clip_test"](fake_self, lq, g)
    finally:
        torch.Tensor.cuda = saved
    save(name, dict(kind="clip_real", c=c, hidden=list(hidden), h=h, w=w, scale=scale, tile=tile, overlap=overlap,
                    seed=seed, eval_bsize=500, non_local=False),
         dict(lq=lq.numpy(), out=out.numpy()))


def swinir_case(name, shapes, seed=0, **kw):
    """The SwinIR trunk (gen_feature) of the reference generator, with the reference's state_dict
    keys and shapes recorded so that checkpoint compatibility is pinned too."""
    g = rh.build_reference_swinir_generator(**kw)
    synth.fill_module(g, seed)
    arrays, tags = {}, []
    with torch.no_grad():
        for (b, h, w) in shapes:
            x = synth.synth_lr_image(b, h, w, seed)
            tag = f"{b}x{h}x{w}"
            arrays[f"x_{tag}"] = x.numpy()
            arrays[f"feat_{tag}"] = g.gen_feature(x)[0].numpy()
            tags.append(tag)
    keys = {k: list(v.shape) for k, v in g.state_dict().items()}
    save(name, dict(kind="swinir", seed=seed, tags=tags, cfg=kw, state_keys=keys), arrays)


def train_case(name, c, hidden, b, h, w, nq, local_size=2, seed=0):
    """Gradients of the training forward (ciaosr.py:88-95: generator(lq, coord, cell) with test_mode=False -> L1
    loss) with respect to the feature map and every head parameter, from the reference's own autograd graph."""
    g = rh.build_reference_generator("edsr", c, hidden, num_blocks=1, eval_bsize=None, local_size=local_size)
    synth.fill_module(g, seed)
    g.train()
    feature = synth.synth_feature(b, c, h, w, seed).requires_grad_(True)
    x_lr = synth.synth_lr_image(b, h, w, seed)
    g.gen_feature = lambda _x: [feature]
    rs = np.random.RandomState(seed + 23)
    coord = torch.from_numpy(rs.uniform(-1, 1, size=(b, nq, 2)).astype(np.float32))
    sc = rs.uniform(1.0, 4.0, size=(b, 1, 1)).astype(np.float32)
    cell = torch.from_numpy(np.broadcast_to(np.concatenate([2.0 / (h * sc), 2.0 / (w * sc)], axis=2), (b, nq, 2)).copy())
    gt = torch.from_numpy(rs.uniform(-0.5, 0.5, size=(b, nq, 3)).astype(np.float32))
    pred = g(x_lr, coord, cell)
    loss = (pred - gt).abs().mean()
    loss.backward()
    arrays = dict(coord=coord.numpy(), cell=cell.numpy(), gt=gt.numpy(), pred=pred.detach().numpy(),
                  loss=np.array([float(loss)], dtype=np.float32), grad_feature=feature.grad.numpy())
    for k, v in g.named_parameters():
        if k.startswith(("imnet_", "cs_attn.")) and v.grad is not None:
            arrays["grad__" + k] = v.grad.numpy()
    save(name, dict(kind="train", c=c, hidden=list(hidden), b=b, h=h, w=w, nq=nq, local_size=local_size, seed=seed,
                    non_local=True, eval_bsize=None), arrays)


def save(name, meta, arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                        **{k: np.ascontiguousarray(v, dtype=np.float32) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  {meta}")


def main(only=()):
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if only:                                  # python -m oracle.make_golden train_small : just the named late additions
        train_case("train_small", 16, (32, 32), 2, 7, 6, 60, seed=16)
        return
    # generic small shapes (SIMT engine): odd H (reflect pad), non-square, several chunks
    head_case("head_small", 8, (32, 32), 2, 7, 6, [2, 3], eval_bsize=100, random_q=150)
    head_case("head_frac", 8, (16,), 1, 6, 5, [2.5, 1.7], eval_bsize=None, seed=1)
    head_case("head_ls1", 8, (16, 16), 1, 6, 6, [2], eval_bsize=50, local_size=1, seed=2)
    head_case("head_ls3", 8, (16, 16), 1, 6, 6, [3], eval_bsize=None, local_size=3, seed=3)
    head_case("head_nonl0", 16, (32, 32, 32), 2, 6, 8, [2, 4], eval_bsize=300, non_local=False, seed=4)
    # the real head dimensions (tcgen05 engine): C=64, hidden 256x4
    head_case("head_c64", 64, (256, 256, 256, 256), 2, 12, 10, [2, 4], eval_bsize=700, seed=5,
              random_q=300)
    head_case("head_c64_nonl0", 64, (256, 256, 256, 256), 1, 10, 12, [3], eval_bsize=None,
              non_local=False, seed=6)
    # SwinIR-sized head (configs 001-swinir / 002: C = 180, 9C is not a multiple of 64 or 128)
    head_case("head_c180", 180, (256, 256, 256, 256), 1, 8, 10, [3], eval_bsize=1000, seed=12)
    head_case("head_c180_nonl0", 180, (256, 256, 256, 256), 1, 6, 8, [4], eval_bsize=None,
              non_local=False, seed=13)
    # cross-scale attention alone
    csattn_case("csattn_c64", 64, 1, 24, 20, seed=7)
    csattn_case("csattn_c180", 180, 1, 12, 16, seed=14)
    csattn_case("csattn_odd", 16, 2, 9, 11, seed=8)
    # tiled inference through the restorer
    clip_case("clip_small", 16, (32, 32), 40, 36, 2, 24, 8, seed=9)
    real_clip_case("clip_real_small", 16, (32, 32), 30, 44, 4, 16, 4, seed=15)
    # training forward + backward through the reference's autograd graph
    train_case("train_small", 16, (32, 32), 2, 7, 6, 60, seed=16)
    # SwinIR trunk: window-multiple size (buffered shift mask), padded sizes (reflect pad + recomputed mask)
    swinir_case("swinir_trunk", [(1, 8, 8), (2, 10, 13), (1, 16, 12)], seed=10,
                embed_dim=24, depths=(2, 2), num_heads=(2, 2), window_size=4, img_size=8, mlp_ratio=2)


if __name__ == "__main__":
    main(tuple(sys.argv[1:]))
