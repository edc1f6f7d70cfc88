"""SwinIR encoder trunk (host PyTorch, per BASELINE.json north_star: "host code stays
Python/PyTorch for the EDSR/RDN/SwinIR encoder").

Mirrors what ``LocalImplicitSRSWINIR`` hoists from the reference's ``SwinIR``
(mmedited/models/backbones/sr_backbones/swinir_net.py:619-760 via ciaosr_net.py:456-470):
``conv_first, patch_embed, pos_drop, layers, norm, patch_unembed, conv_after_body`` and
``embed_dim``, with the reference's parameter / buffer names
(``layers.N.residual_group.blocks.M.{norm1, attn.{relative_position_bias_table,
relative_position_index, qkv, proj}, norm2, mlp.{fc1, fc2}, attn_mask}``, ``layers.N.conv``,
``patch_embed.norm``, ``norm``), so released checkpoints map onto ``generator.*`` unchanged.
The reconstruction tail (``conv_before_upsample / upsample / conv_last``, swinir_net.py:737-762)
is dropped by the generator (``del self.encoder``, ciaosr_net.py:472) and not built here;
its keywords are accepted and ignored.

Written for the GPU the rest of the package targets rather than transcribed: windows are
gathered with one reshape/permute, attention is ``F.scaled_dot_product_attention`` with the
relative-position bias and the shifted-window mask folded into one additive mask (cached
per input size and device), stochastic depth is an inference no-op.  Unlike the reference
(``.cuda()`` calls in its constructor, swinir_net.py:684,723,725) it builds on any device.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


# ---- encoder fast path: the trunk's Linear layers on the tensor cores -------------------------
# The trunk's FLOPs are its Linear layers (qkv, proj, fc1, fc2; K = 180 / 360).  Parity with the fp32 reference
# rules out TF32 (features ~1e-3 off), and cuBLAS' fp32 CUDA-core sgemm runs these shapes at ~14 TFLOP/s (12 of the
# 29 ms a 128x128 tile spends in the trunk).  Inside `native_linear(True)` they go through the library's
# `ciaosr_linear_forward` instead: the fp16 hi/lo-split tcgen05 GEMM of the head (fp32-grade) with bias and the
# exact GELU fused in its epilogue.  Everything else of the trunk stays plain PyTorch.
_NATIVE = False


class native_linear:
    """Context manager: route this module's nn.Linear layers through the native tensor-core Linear."""

    def __init__(self, enabled):
        self.enabled = bool(enabled)

    def __enter__(self):
        global _NATIVE
        self.prev, _NATIVE = _NATIVE, self.enabled

    def __exit__(self, *exc):
        global _NATIVE
        _NATIVE = self.prev


def _norm(x, ln):
    """`ln(x)` for an affine nn.LayerNorm over the last dimension."""
    if (_NATIVE and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and ln.elementwise_affine
            and len(ln.normalized_shape) == 1 and ln.normalized_shape[0] <= 512):
        from . import native
        return native.layernorm(x, ln)
    return ln(x)


def _conv3x3_tokens(tokens, x_size, conv, residual=None):
    """`conv` (an nn.Conv2d 3x3, stride 1, pad 1) applied to a token tensor [B, HW, C] seen as the NHWC map it is,
    + residual tokens; returns tokens, or None when the native path does not apply."""
    if _NATIVE and tokens.is_cuda and tokens.dtype == torch.float32 and not torch.is_grad_enabled():
        from . import native
        plan = native.conv3x3_plan_for(conv)
        if plan is not None:
            b, n, c = tokens.shape
            return plan.forward(tokens.view(b, x_size[0], x_size[1], c), residual=residual).view(b, n, -1)
    return None


def _linear(x, lin, gelu=False, residual=None):
    """`lin(x)` (optionally followed by the exact GELU, optionally + residual) for an nn.Linear."""
    w = lin.weight
    if _NATIVE and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled():
        from . import native
        if native.LinearPlan.supports(w):
            key = (w.data_ptr(), w._version, None if lin.bias is None else (lin.bias.data_ptr(), lin.bias._version))
            mc = native.module_cache(lin)             # not in lin.__dict__: keeps the module deepcopy- / pickle-able
            cache = mc.get("linear_plan")
            if cache is None or cache[0] != key:
                cache = mc["linear_plan"] = (key, native.LinearPlan(w, lin.bias))
            return cache[1].forward(x, gelu=gelu, residual=residual)
    y = lin(x)
    y = F.gelu(y) if gelu else y
    return y if residual is None else residual + y


class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath); identity in eval mode."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


class Mlp(nn.Module):
    """swinir_net.py:15-31."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        if isinstance(self.act, nn.GELU) and getattr(self.act, "approximate", "none") == "none":
            return self.drop(_linear(self.drop(_linear(x, self.fc1, gelu=True)), self.fc2))
        return self.drop(_linear(self.drop(self.act(_linear(x, self.fc1))), self.fc2))


def relative_position_index(window):
    """[N, N] index into the (2Wh-1)(2Ww-1) bias table (swinir_net.py:93-104)."""
    wh, ww = window
    ys, xs = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
    pos = torch.stack([ys.reshape(-1), xs.reshape(-1)])              # [2, N]
    rel = pos[:, :, None] - pos[:, None, :]                          # [2, N, N]
    return (rel[0] + wh - 1) * (2 * ww - 1) + (rel[1] + ww - 1)


def shifted_window_mask(size, window, shift):
    """[nW, N, N] additive mask (0 / -100) of SW-MSA for an H x W map (swinir_net.py:217-238)."""
    h, w = size
    region = torch.zeros(h, w)
    bounds_h = (0, h - window, h - shift, h)
    bounds_w = (0, w - window, w - shift, w)
    label = 0
    for i in range(3):
        for j in range(3):
            region[bounds_h[i]:bounds_h[i + 1], bounds_w[j]:bounds_w[j + 1]] = label
            label += 1
    win = region.view(h // window, window, w // window, window).permute(0, 2, 1, 3).reshape(-1, window * window)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class WindowAttention(nn.Module):
    """W-MSA / SW-MSA with relative position bias (swinir_net.py:66-146)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, _pair(window_size), num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * wh - 1) * (2 * ww - 1), num_heads))
        self.register_buffer("relative_position_index", relative_position_index(self.window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def bias(self):
        n = self.window_size[0] * self.window_size[1]
        return self.relative_position_bias_table[self.relative_position_index.reshape(-1)] \
            .view(n, n, self.num_heads).permute(2, 0, 1)                # [nH, N, N]

    def forward(self, x, mask=None):
        """x [B * nW, N, C]; mask [nW, N, N] additive or None."""
        b_, n, c = x.shape
        qkv = _linear(x, self.qkv).view(b_, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        add = self.bias().unsqueeze(0)                                  # [1, nH, N, N]
        if mask is not None:
            nw = mask.shape[0]
            add = (add.unsqueeze(1) + mask.view(1, nw, 1, n, n)).expand(b_ // nw, nw, self.num_heads, n, n) \
                .reshape(b_, self.num_heads, n, n)
        out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=add.to(x.dtype),
                                             dropout_p=self.attn_drop.p if self.training else 0.0,
                                             scale=self.scale)
        return self.proj_drop(_linear(out.transpose(1, 2).reshape(b_, n, c), self.proj))


class SwinTransformerBlock(nn.Module):
    """swinir_net.py:165-280."""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0,
                 qkv_bias=True, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, tuple(input_resolution), num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:        # window larger than the map: one window, no shift
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, self.window_size, num_heads, qkv_bias, qk_scale, attn_drop, drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.register_buffer("attn_mask", shifted_window_mask(self.input_resolution, self.window_size,
                                                              self.shift_size) if self.shift_size > 0 else None)
        self._masks = {}

    def _mask_for(self, x_size, device):
        if self.shift_size == 0:
            return None
        if tuple(x_size) == self.input_resolution:
            return self.attn_mask
        key = (tuple(x_size), str(device))
        if key not in self._masks:
            if len(self._masks) > 16:
                self._masks.clear()
            self._masks[key] = shifted_window_mask(x_size, self.window_size, self.shift_size).to(device)
        return self._masks[key]

    def forward(self, x, x_size):
        h, w = x_size
        b, _, c = x.shape
        ws, sh = self.window_size, self.shift_size
        if _NATIVE and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled():
            from . import native
            a = self.attn
            m = self.mlp
            plans = [native.linear_plan_for(l) for l in (a.qkv, a.proj, m.fc1, m.fc2)]
            exact_gelu = isinstance(m.act, nn.GELU) and getattr(m.act, "approximate", "none") == "none"
            if (native.window_attention_supported(c, a.num_heads, ws) and all(p is not None for p in plans)
                    and exact_gelu and (c // a.num_heads) % 2 == 0 and self.norm1.elementwise_affine
                    and self.norm2.elementwise_affine and c <= 512):
                # Encoder fast path: the whole block in 7 native kernels.  Activations travel between them in the fp16
                # hi / lo split form the tensor-core GEMMs compute in (native.SplitTensor): LayerNorm, the attention
                # kernel and fc1's epilogue WRITE that form, so every Linear's operand tiles are TMA-loaded instead of
                # being gathered and re-split from fp32 by the GEMM's row threads (the bound of these short-K GEMMs).
                # The attention kernel reads the qkv Linear's output in natural token order: shift, window partition,
                # relative position bias, SW-MSA mask and their inverses are index arithmetic (csrc/swin_attn.cu).
                qkv_p, proj_p, fc1_p, fc2_p = plans
                qkv = qkv_p.forward_split(native.layernorm_split(x, self.norm1), out_shape=(b, h * w))
                o = native.window_attention(qkv, a.relative_position_bias_table, h, w, a.num_heads, ws, sh, a.scale,
                                            split_out=True)
                x = proj_p.forward_split(o, residual=x, out_shape=(b, h * w))          # x + proj(o)
                g = fc1_p.forward_split(native.layernorm_split(x, self.norm2), gelu=True, split_out=True)
                return fc2_p.forward_split(g, residual=x, out_shape=(b, h * w))        # x + fc2(gelu(fc1(.)))
        y = self.norm1(x).view(b, h, w, c)
        if sh > 0:
            y = torch.roll(y, shifts=(-sh, -sh), dims=(1, 2))
        win = y.view(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, c)
        win = self.attn(win, self._mask_for(x_size, x.device))
        y = win.view(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
        if sh > 0:
            y = torch.roll(y, shifts=(sh, sh), dims=(1, 2))
        x = x + self.drop_path(y.view(b, h * w, c))
        return x + self.drop_path(self.mlp(self.norm2(x)))


class BasicLayer(nn.Module):
    """swinir_net.py:350-406 (no downsampling in SwinIR)."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm,
                 downsample=None, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, input_resolution, num_heads, window_size,
                                 0 if i % 2 == 0 else window_size // 2, mlp_ratio, qkv_bias, qk_scale, drop,
                                 attn_drop, drop_path[i] if isinstance(drop_path, (list, tuple)) else drop_path,
                                 norm_layer=norm_layer)
            for i in range(depth)])
        self.downsample = None

    def forward(self, x, x_size):
        for blk in self.blocks:
            x = blk(x, x_size)
        return x


class PatchEmbed(nn.Module):
    """[B, C, H, W] -> [B, HW, C] (+ LayerNorm) (swinir_net.py:496-529)."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.img_size, self.patch_size = _pair(img_size), _pair(patch_size)
        self.patches_resolution = [self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        x = x.flatten(2).transpose(1, 2)
        return _norm(x, self.norm) if self.norm is not None else x


class PatchUnEmbed(nn.Module):
    """[B, HW, C] -> [B, C, H, W] (swinir_net.py:539-566)."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.img_size, self.patch_size = _pair(img_size), _pair(patch_size)
        self.patches_resolution = [self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim

    def forward(self, x, x_size):
        return x.transpose(1, 2).reshape(x.shape[0], self.embed_dim, x_size[0], x_size[1])


def _resi_conv(dim, kind):
    if kind == "1conv":
        return nn.Conv2d(dim, dim, 3, 1, 1)
    if kind == "3conv":
        return nn.Sequential(nn.Conv2d(dim, dim // 4, 3, 1, 1), nn.LeakyReLU(0.2, inplace=True),
                             nn.Conv2d(dim // 4, dim // 4, 1, 1, 0), nn.LeakyReLU(0.2, inplace=True),
                             nn.Conv2d(dim // 4, dim, 3, 1, 1))
    raise ValueError(f"resi_connection must be '1conv' or '3conv', got {kind!r}")


class RSTB(nn.Module):
    """Residual Swin Transformer block (swinir_net.py:420-483)."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None,
                 use_checkpoint=False, img_size=224, patch_size=4, resi_connection="1conv"):
        super().__init__()
        self.dim, self.input_resolution = dim, input_resolution
        self.residual_group = BasicLayer(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias,
                                         qk_scale, drop, attn_drop, drop_path, norm_layer)
        self.conv = _resi_conv(dim, resi_connection)
        self.patch_embed = PatchEmbed(img_size, patch_size, 0, dim, None)
        self.patch_unembed = PatchUnEmbed(img_size, patch_size, 0, dim, None)

    def forward(self, x, x_size):
        t = self.residual_group(x, x_size)
        if isinstance(self.conv, nn.Conv2d):
            y = _conv3x3_tokens(t, x_size, self.conv, residual=x)     # tokens are an NHWC map: no (un)embed transposes
            if y is not None:
                return y
        y = self.patch_unembed(t, x_size)
        return self.patch_embed(self.conv(y)) + x


class SwinIR(nn.Module):
    """Trunk of SwinIR with the reference's constructor keywords (swinir_net.py:647-653)."""

    def __init__(self, img_size=64, patch_size=1, in_chans=3, embed_dim=96, depths=(6, 6, 6, 6),
                 num_heads=(6, 6, 6, 6), window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False,
                 patch_norm=True, use_checkpoint=False, upscale=2, img_range=1.0, upsampler="",
                 resi_connection="1conv", **kwargs):
        super().__init__()
        self.img_range, self.upscale, self.upsampler, self.window_size = img_range, upscale, upsampler, window_size
        self.embed_dim = self.num_features = embed_dim
        self.num_layers, self.ape, self.patch_norm, self.mlp_ratio = len(depths), ape, patch_norm, mlp_ratio
        self.conv_first = nn.Conv2d(in_chans, embed_dim, 3, 1, 1)
        self.patch_embed = PatchEmbed(img_size, patch_size, embed_dim, embed_dim, norm_layer if patch_norm else None)
        self.patches_resolution = self.patch_embed.patches_resolution
        self.patch_unembed = PatchUnEmbed(img_size, patch_size, embed_dim, embed_dim, None)
        if ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=0.02)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [float(v) for v in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList([
            RSTB(embed_dim, tuple(self.patches_resolution), depths[i], num_heads[i], window_size, mlp_ratio, qkv_bias,
                 qk_scale, drop_rate, attn_drop_rate, dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer, None,
                 use_checkpoint, img_size, patch_size, resi_connection)
            for i in range(len(depths))])
        self.norm = norm_layer(embed_dim)
        self.conv_after_body = _resi_conv(embed_dim, resi_connection)

    def init_weights(self, pretrained=None, strict=True):
        """swinir_net.py:773-791."""
        from .builder import load_checkpoint
        if isinstance(pretrained, str):
            load_checkpoint(self, pretrained, strict=strict)
        elif pretrained is None:
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    nn.init.trunc_normal_(m.weight, std=0.02)
                    if m.bias is not None:
                        nn.init.zeros_(m.bias)
                elif isinstance(m, nn.LayerNorm):
                    nn.init.zeros_(m.bias)
                    nn.init.ones_(m.weight)
        else:
            raise TypeError(f'"pretrained" must be a str or None. But received {type(pretrained)}.')
