"""Training support of the head: gradients for ``generator(lq, coord, cell)`` as
``CiaoSR.train_step`` calls it (mmedited/models/restorers/ciaosr.py:60-109 -> ciaosr_net.py:88-110, 113-224).

The forward value always comes from the native kernels (``HeadFunction.forward`` -> ``HeadPlan.query_rgb``; there is
no CPU / PyTorch forward path).  The BACKWARD is a recompute: the functions below restate the head with
differentiable torch ops on explicit gather indices (no ``unfold`` of the whole map, no ``grid_sample``), are run
under ``torch.enable_grad()`` on the saved inputs, and ``torch.autograd.grad`` pulls the incoming gradient back to the
feature map and to every head parameter.  It is allowed to be slow; it has to be right: pinned on gradients the
unmodified reference produced (tests/golden/train_small.npz, oracle/make_golden.py).

Parameters travel as a flat ``{state_dict key: tensor}`` dict (``imnet_k.layers.0.weight``, ``cs_attn.down.bias`` ...).
"""
import torch
import torch.nn.functional as F


def _mlp(x, p, name):
    """MLPRefiner (mlp_refiner.py:65-102): Linear + ReLU ... Linear over keys ``name.layers.{0,2,4,..}``."""
    i = 0
    while f"{name}.layers.{i}.weight" in p:
        x = F.linear(x, p[f"{name}.layers.{i}.weight"], p[f"{name}.layers.{i}.bias"])
        i += 2
        if f"{name}.layers.{i}.weight" in p:
            x = torch.relu(x)
    return x


def _nearest(c, n):
    """Source index of grid_sample(nearest, align_corners=False): round-half-even of ((c + 1) n - 1) / 2, fp32."""
    return torch.round(((c + 1.0) * n - 1.0) / 2.0).long()


def _pixel_centres(n, like):
    r = 1.0 / n
    return ((-1.0 + r) + (2.0 * r) * torch.arange(n, dtype=torch.float32, device=like.device)).to(like.dtype)


def _gather_unfolded(fpad, nl, iy, ix):
    """Rows of ``unfold(feature, 3, padding=1)`` (channel order c*9 + tap) at pixels (iy, ix), zero outside the map,
    optionally followed by the non-local channels of the same pixel.  fpad = feature zero-padded by 1: [B,C,H+2,W+2]."""
    b, c, hp, wp = fpad.shape
    h, w = hp - 2, wp - 2
    ok = (iy >= 0) & (iy < h) & (ix >= 0) & (ix < w)
    iyc, ixc = iy.clamp(0, h - 1), ix.clamp(0, w - 1)
    bi = torch.arange(b, device=fpad.device).view(b, 1).expand_as(iy)
    taps = [fpad[bi, :, iyc + ky, ixc + kx] for ky in range(3) for kx in range(3)]         # 9 x [B,q,C]
    out = torch.stack(taps, dim=-1).reshape(b, iy.shape[1], c * 9)
    if nl is not None:
        out = torch.cat([out, nl[bi, :, iyc, ixc]], dim=-1)
    return out * ok.unsqueeze(-1).to(out.dtype)


def cross_scale_attention(x, p, softmax_scale=10.0, prefix="cs_attn"):
    """CrossScaleAttention.forward for scale [2] (arch_csnln.py:430-532) in attention form, differentiable:
    softmax over the 3x3 patches of the half-resolution map, 6x6 stride-2 value patches folded back and
    reduced by the 3x3 stride-2 `down` convolution, / 6."""
    b, c, h, w = x.shape
    xp = F.pad(x, (0, w % 2, 0, h % 2), mode="reflect") if (h % 2 or w % 2) else x
    hp, wp = xp.shape[-2:]

    def block(t, name):
        y = F.conv2d(t, p[f"{prefix}.{name}.0.weight"], p[f"{prefix}.{name}.0.bias"])
        return F.prelu(y, p[f"{prefix}.{name}.1.weight"])
    emb, mat = block(xp, "conv_assembly"), block(xp, "conv_match_1")
    ref = block(F.avg_pool2d(xp, 2), "conv_match_2")           # bilinear x0.5 of an even-sized map = 2x2 mean
    q = F.unfold(mat, 3, padding=1).transpose(1, 2)                                   # [B, HpWp, 9C/2]
    k = F.unfold(ref, 3, padding=1).transpose(1, 2)                                   # [B, L, 9C/2]
    k = k / torch.maximum(k.norm(dim=2, keepdim=True), p[f"{prefix}.escape_NaN"].view(1, 1, 1))
    prob = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * softmax_scale, dim=2)      # [B, HpWp, L]
    v = F.unfold(emb, 6, stride=2, padding=2)                                         # [B, 36C, L]
    o = torch.bmm(prob, v.transpose(1, 2)).transpose(1, 2)                            # [B, 36C, HpWp]
    canvas = F.fold(o, (2 * hp, 2 * wp), 6, stride=2, padding=2)                      # conv_transpose2d(stride 2, pad 2)
    y = F.conv2d(canvas, p[f"{prefix}.down.weight"], p[f"{prefix}.down.bias"], stride=2, padding=1) / 6.0
    return y[:, :, :h, :w]


def query_rgb(feature, coord, cell, p, local_size=2, non_local_attn=True, softmax_scale=1.0, cs_softmax_scale=10.0):
    """ciaosr_net.py:113-224 for ``feat_unfold=True``: feature [B,C,H,W], coord / cell [B,q,2] -> [B,q,3]."""
    b, c, h, w = feature.shape
    q = coord.shape[1]
    nl = cross_scale_attention(feature, p, cs_softmax_scale) if non_local_attn else None
    fpad = F.pad(feature, (1, 1, 1, 1))
    query = _gather_unfolded(fpad, None, _nearest(coord[..., 0], h), _nearest(coord[..., 1], w))
    seq_y, seq_x = _pixel_centres(h, coord), _pixel_centres(w, coord)
    offs = [0] if local_size == 1 else list(range(-1, 2, 4 - local_size))
    tx = ((h - 1) / (1 - cell[:, 0, 0])).view(b, 1)
    ty = ((w - 1) / (1 - cell[:, 0, 1])).view(b, 1)
    sc = torch.stack([cell[..., 0] * h, cell[..., 1] * w], dim=-1)
    logits, values = [], []
    for vx in offs:
        for vy in offs:
            cy, cx = coord[..., 0], coord[..., 1]
            if vx != 0:
                cy = cy + (vx / abs(vx) * ((2 * abs(vx) - 1) / tx) + 1e-6)
            if vy != 0:
                cx = cx + (vy / abs(vy) * ((2 * abs(vy) - 1) / ty) + 1e-6)
            cy, cx = cy.clamp(-1 + 1e-6, 1 - 1e-6), cx.clamp(-1 + 1e-6, 1 - 1e-6)
            iy, ix = _nearest(cy, h), _nearest(cx, w)
            val = _gather_unfolded(fpad, nl, iy, ix)                                   # [B,q,9C(+Cn)]
            key = val[..., :9 * c]
            rel = torch.stack([(coord[..., 0] - seq_y[iy.clamp(0, h - 1)]) * h,
                               (coord[..., 1] - seq_x[ix.clamp(0, w - 1)]) * w], dim=-1)
            pk = key * _mlp(torch.cat([key, rel, sc], dim=-1), p, "imnet_k")
            pv = val * _mlp(torch.cat([val, rel, sc], dim=-1), p, "imnet_v")
            logits.append((query * pk).sum(-1))
            values.append(pv)
    attn = torch.softmax(torch.stack(logits, dim=-1) / softmax_scale, dim=-1)         # [B,q,n]
    x = (attn.unsqueeze(-1) * torch.stack(values, dim=-2)).sum(-2)                    # [B,q,Dv]
    return _mlp(x, p, "imnet_q").view(b, q, -1)


class HeadFunction(torch.autograd.Function):
    """forward: the native head (incl. the bilinear residual when `lr_image` is given); backward: recompute above.
    apply(plan, hyper, feature, coord, cell, lr_image, keys, *param_tensors)."""

    @staticmethod
    def forward(ctx, plan, hyper, feature, coord, cell, lr_image, keys, *params):
        ctx.hyper, ctx.keys = hyper, keys
        ctx.save_for_backward(feature, coord, cell, *params)
        with torch.no_grad():
            return plan.query_rgb(feature.detach(), coord, cell, lr_image=lr_image, eval_bsize=None,
                                  engine=hyper["engine"])

    @staticmethod
    def backward(ctx, grad_out):
        feature, coord, cell, *params = ctx.saved_tensors
        need_f = ctx.needs_input_grad[2]
        with torch.enable_grad():
            f = feature.detach().requires_grad_(need_f)
            ps = [t.detach().requires_grad_(True) for t in params]
            h = ctx.hyper
            out = query_rgb(f, coord, cell, dict(zip(ctx.keys, ps)), local_size=h["local_size"],
                            non_local_attn=h["non_local_attn"], softmax_scale=h["softmax_scale"],
                            cs_softmax_scale=h["cs_softmax_scale"])
            wanted = ([f] if need_f else []) + [t for t in ps if t.is_floating_point()]
            grads = list(torch.autograd.grad(out, wanted, grad_out.contiguous(), allow_unused=True))
        gf = grads.pop(0) if need_f else None
        gp = [grads.pop(0) if t.is_floating_point() else None for t in ps]
        return (None, None, gf, None, None, None, None, *gp)
