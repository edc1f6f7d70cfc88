"""CPU oracle for the CiaoSR implicit attention-in-attention head.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker / the timed
CPU baseline.  The product path (``ciaosr_b200``) never imports it and has no
CPU fallback.

What this is: an independent restatement, in explicit index arithmetic on
fp32 torch-CPU tensors, of the algorithm the reference implements with
``F.unfold`` / ``F.grid_sample`` / ``conv2d`` / ``conv_transpose2d``:

  ================  ==========================================================
  function here     reference lines it restates (paths under /root/reference)
  ================  ==========================================================
  make_coord        mmedit 0.11 ``make_coord`` (not vendored); call sites
                    mmedited/models/backbones/sr_backbones/ciaosr_net.py:148,
                    mmedited/models/restorers/ciaosr.py:240
  mlp               mmedited/models/components/refiners/mlp_refiner.py:65-102
  cross_scale_attention
                    mmedited/models/common/arch_csnln.py:430-532 (+ helpers
                    :32-87)
  nearest_index     ATen grid_sampler_2d(nearest, align_corners=False) as used
                    at ciaosr_net.py:145,176,178,182
  unfold3x3_at      F.unfold(feature, 3, padding=1) at ciaosr_net.py:131-139
                    evaluated only at the gathered pixels
  query_rgb         ciaosr_net.py:113-224
  batched_predict   ciaosr_net.py:226-248
  bilinear_border   grid_sample(bilinear, border) at ciaosr_net.py:107-108
  head_forward      ciaosr_net.py:88-110 minus the encoder (takes `feature`)
  clip_test         mmedited/models/restorers/ciaosr.py:218-258
  ================  ==========================================================

Pinning: the reference ships NO golden vectors or tests.
This oracle is pinned against outputs of the reference itself, executed in
the build container through ``oracle/ref_harness.py`` and frozen by
``oracle/make_golden.py`` into ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks the oracle against those files (and,
when ``/root/reference`` is mounted, against the live reference).

Weights are passed as a flat ``dict[str, Tensor]`` using the reference's
state_dict key names relative to the generator (``imnet_k.layers.0.weight``,
``cs_attn.conv_match_1.0.weight`` ...).
"""
import math

import torch

F32 = torch.float32


# ----------------------------------------------------------------------------
# coordinates
# ----------------------------------------------------------------------------
def make_coord(shape, flatten=True):
    """Pixel-centre coordinates in [-1, 1], (y, x) order.

    seq_i = (-1 + 1/n) + (2/n) * i, the two constants formed in double and
    rounded to fp32 where they meet the fp32 ``arange``.
    """
    seqs = []
    for n in shape:
        r = 1.0 / n
        seqs.append((-1.0 + r) + (2.0 * r) * torch.arange(n, dtype=F32))
    grid = torch.stack(torch.meshgrid(*seqs, indexing="ij"), dim=-1)
    return grid.reshape(-1, len(shape)) if flatten else grid


def cell_for(target_hw, q):
    """cell tensor the callers build: (2/Ht, 2/Wt) per query (ciaosr.py:241-243)."""
    c = torch.ones(q, 2, dtype=F32)
    c[:, 0] *= 2 / target_hw[0]
    c[:, 1] *= 2 / target_hw[1]
    return c


def nearest_index(c, n):
    """Nearest source index for a normalised coordinate (align_corners=False).

    u = ((c + 1) * n - 1) / 2 in fp32, rounded half-to-even; indices outside
    [0, n) select the zero padding.
    """
    u = ((c + 1.0) * float(n) - 1.0) / 2.0
    return torch.round(u).to(torch.int64)


# ----------------------------------------------------------------------------
# MLP
# ----------------------------------------------------------------------------
def mlp_layers(w, prefix):
    """[(weight[out,in], bias[out])...] for keys ``prefix.layers.{0,2,4,...}``."""
    out, i = [], 0
    while f"{prefix}.layers.{i}.weight" in w:
        out.append((w[f"{prefix}.layers.{i}.weight"], w[f"{prefix}.layers.{i}.bias"]))
        i += 2
    return out


def mlp(x, layers):
    """Linear+ReLU for every layer but the last, which is a bare Linear."""
    for li, (wt, b) in enumerate(layers):
        x = torch.addmm(b, x, wt.t())
        if li + 1 < len(layers):
            x = torch.relu(x)
    return x


# ----------------------------------------------------------------------------
# cross-scale non-local attention
# ----------------------------------------------------------------------------
def _conv1x1_prelu(x, wt, b, slope):
    # x [C,H,W] -> [Co,H,W]
    c, h, w_ = x.shape
    y = torch.addmm(b[:, None], wt.reshape(wt.shape[0], c), x.reshape(c, h * w_))
    y = torch.where(y >= 0, y, y * slope.reshape(-1, 1))
    return y.reshape(-1, h, w_)


def _reflect_pad_br(x, pad_h, pad_w):
    # reflect padding on the bottom / right only (arch_csnln.py:444-449)
    c, h, w_ = x.shape
    if pad_h:
        idx = torch.arange(h - 2, h - 2 - pad_h, -1)
        x = torch.cat([x, x[:, idx, :]], dim=1)
    if pad_w:
        idx = torch.arange(w_ - 2, w_ - 2 - pad_w, -1)
        x = torch.cat([x, x[:, :, idx]], dim=2)
    return x


def _bilinear_down(x, s):
    # F.interpolate(scale_factor=1/s, bilinear, align_corners=False) on a size
    # divisible by s: src = (o + 0.5) * s - 0.5
    c, h, w_ = x.shape

    def taps(n_out, n_in):
        src = (torch.arange(n_out, dtype=F32) + 0.5) * float(s) - 0.5
        src = src.clamp(min=0)
        i0 = src.floor().to(torch.int64)
        i1 = torch.clamp(i0 + 1, max=n_in - 1)
        l1 = src - i0.to(F32)
        return i0, i1, 1.0 - l1, l1

    y0, y1, wy0, wy1 = taps(h // s, h)
    x0, x1, wx0, wx1 = taps(w_ // s, w_)
    rows = x[:, y0, :] * wy0[None, :, None] + x[:, y1, :] * wy1[None, :, None]
    return rows[:, :, x0] * wx0[None, None, :] + rows[:, :, x1] * wx1[None, None, :]


def _patches(x, k, stride, pad):
    """[C,H,W] -> [L, C, k, k] patches, zero padded by `pad` on every side."""
    c, h, w_ = x.shape
    xp = torch.zeros(c, h + 2 * pad, w_ + 2 * pad, dtype=x.dtype)
    xp[:, pad:pad + h, pad:pad + w_] = x
    ny = (h + 2 * pad - k) // stride + 1
    nx = (w_ + 2 * pad - k) // stride + 1
    oy = (torch.arange(ny) * stride)[:, None] + torch.arange(k)[None, :]   # [ny,k]
    ox = (torch.arange(nx) * stride)[:, None] + torch.arange(k)[None, :]   # [nx,k]
    p = xp[:, oy[:, None, :, None], ox[None, :, None, :]]                  # [C,ny,nx,k,k]
    return p.permute(1, 2, 0, 3, 4).reshape(ny * nx, c, k, k), ny, nx


def cross_scale_attention_one(x, w, scales=(2,), softmax_scale=10.0, prefix="cs_attn"):
    """One image: x [C,H,W] -> [C*len(scales), H, W].

    Attention form of the reference's conv2d / softmax / conv_transpose2d:
      Q  = 3x3 zero-padded patches of match_1(x)                 [HW, 9C/2]
      K^ = 3x3 patches of match_2(down_s(x)), L2-normalised with
           floor 1e-4                                             [L, 9C/2]
      V  = (3s)x(3s) stride-s patches of assembly(x), pad s       [L, C(3s)^2]
      P  = softmax_L(softmax_scale * Q K^T)
      canvas[c, s*y - s + i, s*x - s + j] += (P V)[(y,x), c, i, j]
      out = conv3x3_stride_s(canvas) / 6, cropped to HxW
    """
    c, h, w_ = x.shape
    outs = []
    for s in scales:
        pad_h = (s - h % s) % s
        pad_w = (s - w_ % s) % s
        xp = _reflect_pad_br(x, pad_h, pad_w)
        hp, wp = xp.shape[1:]
        emb = _conv1x1_prelu(xp, w[f"{prefix}.conv_assembly.0.weight"],
                             w[f"{prefix}.conv_assembly.0.bias"],
                             w[f"{prefix}.conv_assembly.1.weight"])
        mat = _conv1x1_prelu(xp, w[f"{prefix}.conv_match_1.0.weight"],
                             w[f"{prefix}.conv_match_1.0.bias"],
                             w[f"{prefix}.conv_match_1.1.weight"])
        ref = _conv1x1_prelu(_bilinear_down(xp, s),
                             w[f"{prefix}.conv_match_2.0.weight"],
                             w[f"{prefix}.conv_match_2.0.bias"],
                             w[f"{prefix}.conv_match_2.1.weight"])
        kp, _, _ = _patches(ref, 3, 1, 1)                       # [L, C/2, 3, 3]
        n_l = kp.shape[0]
        kmat = kp.reshape(n_l, -1)
        knorm = torch.sqrt((kmat * kmat).sum(dim=1, keepdim=True))
        kmat = kmat / torch.maximum(knorm, w[f"{prefix}.escape_NaN"].reshape(1, 1))
        qp, _, _ = _patches(mat, 3, 1, 1)                       # [HpWp, C/2, 3, 3]
        qmat = qp.reshape(hp * wp, -1)
        prob = torch.softmax((qmat @ kmat.t()) * softmax_scale, dim=1)   # [HpWp, L]
        vp, _, _ = _patches(emb, 3 * s, s, s)                   # [L, C, 3s, 3s]
        o = (prob @ vp.reshape(n_l, -1)).reshape(hp, wp, c, 3 * s, 3 * s)
        # overlap-add onto the (s*Hp) x (s*Wp) canvas, offset -s
        canvas = torch.zeros(c, s * hp + 2 * s, s * wp + 2 * s, dtype=F32)
        yy = (torch.arange(hp) * s)[:, None] + torch.arange(3 * s)[None, :]     # [hp,3s] (+s shift cancels -s)
        xx = (torch.arange(wp) * s)[:, None] + torch.arange(3 * s)[None, :]
        flat = (yy[:, None, :, None] * canvas.shape[2] + xx[None, :, None, :])  # [hp,wp,3s,3s]
        canvas.view(c, -1).index_add_(
            1, flat.reshape(-1),
            o.permute(2, 0, 1, 3, 4).reshape(c, -1))
        canvas = canvas[:, s:s + s * hp, s:s + s * wp]
        dname = {2: "down", 3: "downx3", 4: "downx4"}[s]
        dp, ny, nx = _patches(canvas, 3, s, 1)                  # [ny*nx, C, 3, 3]
        dw = w[f"{prefix}.{dname}.weight"].reshape(c, -1)
        y = torch.addmm(w[f"{prefix}.{dname}.bias"][None, :], dp.reshape(ny * nx, -1), dw.t())
        y = (y / 6.0).t().reshape(c, ny, nx)
        outs.append(y[:, :h, :w_])
    return torch.cat(outs, dim=0)


def cross_scale_attention(feature, w, scales=(2,), softmax_scale=10.0, prefix="cs_attn"):
    """feature [B,C,H,W] -> [B, C*len(scales), H, W]; per-image like arch_csnln.py:491."""
    return torch.stack([cross_scale_attention_one(feature[b], w, scales, softmax_scale, prefix)
                        for b in range(feature.shape[0])])


# ----------------------------------------------------------------------------
# implicit attention head
# ----------------------------------------------------------------------------
def unfold3x3_at(feat_b, iy, ix):
    """3x3 zero-padded neighbourhood of feat_b [C,H,W] at pixels (iy, ix).

    Returns [n, 9C] with channel index c*9 + ki*3 + kj (F.unfold's order);
    out-of-range centre pixels give zeros (grid_sample zero padding).
    """
    c, h, w_ = feat_b.shape
    n = iy.shape[0]
    out = torch.zeros(n, c, 9, dtype=F32)
    centre_ok = (iy >= 0) & (iy < h) & (ix >= 0) & (ix < w_)
    for t in range(9):
        yy = iy + (t // 3 - 1)
        xx = ix + (t % 3 - 1)
        ok = centre_ok & (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w_)
        v = feat_b[:, yy.clamp(0, h - 1), xx.clamp(0, w_ - 1)]      # [C,n]
        out[:, :, t] = torch.where(ok[None, :], v, torch.zeros((), dtype=F32)).t()
    return out.reshape(n, c * 9)


def _gather_px(fm_b, iy, ix):
    c, h, w_ = fm_b.shape
    ok = (iy >= 0) & (iy < h) & (ix >= 0) & (ix < w_)
    v = fm_b[:, iy.clamp(0, h - 1), ix.clamp(0, w_ - 1)]
    return torch.where(ok[None, :], v, torch.zeros((), dtype=F32)).t()


def neighbour_offsets(local_size):
    if local_size == 1:
        return [(0, 0)]
    step = 4 - local_size
    return [(i, j) for i in range(-1, 2, step) for j in range(-1, 2, step)]


def query_rgb(feature, coord, cell, w, local_size=2, softmax_scale=1.0,
              non_local_attn=True, multi_scale=(2,), nonlocal_feat=None):
    """feature [B,C,H,W], coord/cell [B,q,2] (y,x) -> [B,q,3].

    `nonlocal_feat` lets a caller inject a precomputed cross-scale attention
    map (used to test the two halves independently); by default it is
    recomputed here, exactly where the reference recomputes it.
    """
    b_, c, h, w_ = feature.shape
    q = coord.shape[1]
    if non_local_attn and nonlocal_feat is None:
        nonlocal_feat = cross_scale_attention(feature, w, multi_scale)
    lk, lv, lq = mlp_layers(w, "imnet_k"), mlp_layers(w, "imnet_v"), mlp_layers(w, "imnet_q")
    seq_y = make_coord((h,))[:, 0]
    seq_x = make_coord((w_,))[:, 0]
    lo = torch.tensor(-1 + 1e-6, dtype=F32)
    hi = torch.tensor(1 - 1e-6, dtype=F32)
    eps = 1e-6
    out = torch.empty(b_, q, 3, dtype=F32)
    for b in range(b_):
        cy, cx = coord[b, :, 0], coord[b, :, 1]
        iqy, iqx = nearest_index(cy, h), nearest_index(cx, w_)
        query = unfold3x3_at(feature[b], iqy, iqx)                        # [q, 9C]
        tx = (h - 1) / (1 - cell[b, 0, 0])
        ty = (w_ - 1) / (1 - cell[b, 0, 1])
        scale_ = torch.stack([cell[b, :, 0] * float(h), cell[b, :, 1] * float(w_)], dim=1)
        logits, vals = [], []
        for vx, vy in neighbour_offsets(local_size):
            sy, sx = cy.clone(), cx.clone()
            if vx != 0:
                sy = sy + ((vx / abs(vx)) * ((2 * abs(vx) - 1) / tx) + eps)
            if vy != 0:
                sx = sx + ((vy / abs(vy)) * ((2 * abs(vy) - 1) / ty) + eps)
            sy = torch.minimum(torch.maximum(sy, lo), hi)
            sx = torch.minimum(torch.maximum(sx, lo), hi)
            iy, ix = nearest_index(sy, h), nearest_index(sx, w_)
            key = unfold3x3_at(feature[b], iy, ix)                        # [q, 9C]
            value = key
            if non_local_attn:
                value = torch.cat([key, _gather_px(nonlocal_feat[b], iy, ix)], dim=1)
            ok_y = (iy >= 0) & (iy < h)
            ok_x = (ix >= 0) & (ix < w_)
            ky = torch.where(ok_y & ok_x, seq_y[iy.clamp(0, h - 1)], torch.zeros((), dtype=F32))
            kx = torch.where(ok_y & ok_x, seq_x[ix.clamp(0, w_ - 1)], torch.zeros((), dtype=F32))
            rel = torch.stack([(cy - ky) * float(h), (cx - kx) * float(w_)], dim=1)
            wk = mlp(torch.cat([key, rel, scale_], dim=1), lk)
            wv = mlp(torch.cat([value, rel, scale_], dim=1), lv)
            logits.append((query * (key * wk)).sum(dim=1))
            vals.append(value * wv)
        attn = torch.softmax(torch.stack(logits, dim=1) / softmax_scale, dim=1)   # [q, n]
        x = (attn[:, :, None] * torch.stack(vals, dim=1)).sum(dim=1)              # [q, Dv]
        out[b] = mlp(x, lq)
    return out


def batched_predict(feature, coord, cell, w, eval_bsize, **kw):
    """Query-axis chunking; every chunk recomputes the cross-scale attention."""
    preds, left, n = [], 0, coord.shape[1]
    while left < n:
        right = min(left + eval_bsize, n)
        preds.append(query_rgb(feature, coord[:, left:right], cell[:, left:right], w, **kw))
        left = right
    return torch.cat(preds, dim=1)


def bilinear_border(img, coord):
    """grid_sample(img, coord.flip(-1), bilinear, border, align_corners=False).

    img [B,3,H,W], coord [B,q,2] (y,x) -> [B,q,3].
    """
    b_, ch, h, w_ = img.shape
    out = torch.empty(b_, coord.shape[1], ch, dtype=F32)
    for b in range(b_):
        def axis(cc, n):
            u = ((cc + 1.0) * float(n) - 1.0) / 2.0
            u = torch.clamp(u, 0.0, float(n - 1))
            i0 = torch.floor(u)
            return i0.to(torch.int64), u - i0, (i0 + 1.0) - u

        y0, fy, gy = axis(coord[b, :, 0], h)
        x0, fx, gx = axis(coord[b, :, 1], w_)
        y1, x1 = y0 + 1, x0 + 1

        def tap(yy, xx, wt):
            ok = (yy < h) & (xx < w_)
            v = img[b][:, yy.clamp(max=h - 1), xx.clamp(max=w_ - 1)]
            return torch.where(ok[None, :], v * wt[None, :], torch.zeros((), dtype=F32))

        acc = tap(y0, x0, gx * gy) + tap(y0, x1, fx * gy) + tap(y1, x0, gx * fy) + tap(y1, x1, fx * fy)
        out[b] = acc.t()
    return out


def head_forward(x_lr, feature, coord, cell, w, eval_bsize=None, test_mode=True, **kw):
    """LocalImplicitSRNet.forward given the encoder output `feature`."""
    if eval_bsize is None or not test_mode:
        pred = query_rgb(feature, coord, cell, w, **kw)
    else:
        pred = batched_predict(feature, coord, cell, w, eval_bsize, **kw)
    return pred + bilinear_border(x_lr, coord)


# ----------------------------------------------------------------------------
# tiled inference (the caller of the boundary)
# ----------------------------------------------------------------------------
def tile_origins(n, tile, overlap):
    stride = tile - overlap
    return list(range(0, n - tile, stride)) + [n - tile]


def clip_test(img_lq, model, scale, tile, tile_overlap):
    """Overlap-average tiling; `model(patch, coord, cell)` -> [B, q, 3]."""
    b_, c, h, w_ = img_lq.shape
    tile = min(tile, h, w_)
    sf = scale
    acc = torch.zeros(b_, c, h * sf, w_ * sf, dtype=F32)
    cnt = torch.zeros_like(acc)
    for y0 in tile_origins(h, tile, tile_overlap):
        for x0 in tile_origins(w_, tile, tile_overlap):
            patch = img_lq[..., y0:y0 + tile, x0:x0 + tile]
            th, tw = round(tile * sf), round(tile * sf)
            coord = make_coord((th, tw)).unsqueeze(0).expand(b_, -1, 2)
            cell = cell_for((th, tw), th * tw).unsqueeze(0).expand(b_, -1, 2)
            o = model(patch, coord, cell)
            o = o.reshape(b_, th, tw, 3).permute(0, 3, 1, 2)
            acc[..., y0 * sf:(y0 + tile) * sf, x0 * sf:(x0 + tile) * sf] += o
            cnt[..., y0 * sf:(y0 + tile) * sf, x0 * sf:(x0 + tile) * sf] += 1.0
    return (acc / cnt).reshape(b_, 3, -1).permute(0, 2, 1).contiguous()


# ----------------------------------------------------------------------------
# bookkeeping used by bench.py / DESIGN.md
# ----------------------------------------------------------------------------
def head_flops_per_query(c, hidden=(256, 256, 256, 256), non_local=True, n_scales=1, n_nb=4):
    """Algorithmic FLOPs (2 x MAC) of the head per output pixel, as the
    reference evaluates it (8 865 280 at C=64)."""
    dk = 9 * c
    dv = dk + (c * n_scales if non_local else 0)

    def mac(i, o):
        dims = [i] + list(hidden) + [o]
        return sum(a * b for a, b in zip(dims[:-1], dims[1:]))

    per_nb = mac(dk + 4, dk) + mac(dv + 4, dv)
    return 2 * (n_nb * per_nb + mac(dv, 3))


def cross_scale_flops(c, h, w_, s=2):
    hw, l_ = h * w_, (h // s) * (w_ // s)
    return 2 * hw * l_ * (9 * (c // 2) + c * 9 * s * s)
