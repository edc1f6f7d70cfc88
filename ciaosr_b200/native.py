"""Host-side driver of the C ABI: descriptors, plan caching, workspace, calls.

PyTorch is used here only as plumbing: it owns device memory (plan buffer,
workspace, outputs) and provides the current CUDA stream.  All arithmetic of
the hot path happens inside ``libciaosr_b200.so``.
"""
import ctypes
import weakref

import torch

from . import _lib
from ._lib import CiaoSRNativeError, ENGINES  # noqa: F401  (re-exported)


# Per-module native state (packed plans, CUDA graphs) lives OUTSIDE the nn.Module, keyed weakly on it: the plans hold
# ctypes structs with pointers, which copy.deepcopy / pickle / torch.save(model) cannot pass through.  A copied or
# unpickled module simply starts with an empty cache and rebuilds its plans lazily from its own parameters.
_MODULE_CACHE = weakref.WeakKeyDictionary()


def module_cache(module):
    d = _MODULE_CACHE.get(module)
    if d is None:
        d = _MODULE_CACHE[module] = {}
    return d


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: ciaosr_b200 has no CPU path for the head")
    return t.contiguous()


def _mlp_desc(layers, keep):
    """layers: list of (weight[out,in], bias[out]) tensors."""
    d = _lib.MlpDesc()
    if len(layers) > _lib.MAX_LAYERS:
        raise ValueError(f"MLP with {len(layers)} Linear layers exceeds CIAOSR_MAX_LAYERS")
    d.n_layers = len(layers)
    for i, (w, b) in enumerate(layers):
        w, b = _f32c(w.detach(), "weight"), _f32c(b.detach(), "bias")
        keep += [w, b]
        if i == 0:
            d.dims[0] = w.shape[1]
        elif w.shape[1] != d.dims[i]:
            raise ValueError(f"layer {i} expects {w.shape[1]} inputs, previous layer gives {d.dims[i]}")
        d.dims[i + 1] = w.shape[0]
        d.weight[i] = w.data_ptr()
        d.bias[i] = b.data_ptr()
    return d


class HeadPlan:
    """Weights of one head, packed on the device for the kernels.

    `params` maps the reference's state_dict key names (relative to the
    generator) to tensors, e.g. ``imnet_k.layers.0.weight``,
    ``cs_attn.conv_match_1.1.weight``.
    """

    def __init__(self, params, channels, local_size=2, non_local_attn=True, multi_scale=(2,),
                 softmax_scale=1.0, feat_unfold=True, cs_softmax_scale=10.0):
        lib = _lib.load()
        self._keep = []
        self.device = params["imnet_q.layers.0.weight"].device
        if self.device.type != "cuda":
            raise RuntimeError("HeadPlan needs CUDA parameters: ciaosr_b200 has no CPU path for the head")
        d = _lib.HeadDesc()
        d.abi_version = _lib.ABI_VERSION
        d.channels = int(channels)
        d.feat_unfold = 1 if feat_unfold else 0
        d.local_size = int(local_size)
        d.non_local_attn = 1 if non_local_attn else 0
        d.softmax_scale = float(softmax_scale)
        for name in ("imnet_q", "imnet_k", "imnet_v"):
            layers, i = [], 0
            while f"{name}.layers.{i}.weight" in params:
                layers.append((params[f"{name}.layers.{i}.weight"], params[f"{name}.layers.{i}.bias"]))
                i += 2
            setattr(d, name, _mlp_desc(layers, self._keep))
        if non_local_attn:
            a = d.cs_attn
            a.channels = int(channels)
            a.n_scales = len(multi_scale)
            if len(multi_scale) > _lib.MAX_SCALES:
                raise ValueError("too many scales")
            for i, s in enumerate(multi_scale):
                a.scales[i] = int(s)
            a.softmax_scale = float(cs_softmax_scale)

            def p(key):
                t = _f32c(params[f"cs_attn.{key}"].detach(), key)
                self._keep.append(t)
                return t.data_ptr()

            a.match1_w, a.match1_b, a.match1_slope = p("conv_match_1.0.weight"), p("conv_match_1.0.bias"), p("conv_match_1.1.weight")
            a.match2_w, a.match2_b, a.match2_slope = p("conv_match_2.0.weight"), p("conv_match_2.0.bias"), p("conv_match_2.1.weight")
            a.assembly_w, a.assembly_b, a.assembly_slope = p("conv_assembly.0.weight"), p("conv_assembly.0.bias"), p("conv_assembly.1.weight")
            a.down_w, a.down_b = p("down.weight"), p("down.bias")
            a.escape_nan = p("escape_NaN")
        self.desc = d
        self.channels = int(channels)
        self.n_nonlocal = int(channels) * len(multi_scale) if non_local_attn else 0
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.ciaosr_plan_bytes(ctypes.byref(d), ctypes.byref(nbytes)))
        with torch.cuda.device(self.device):
            self.buf = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=self.device)
            _lib.check(lib.ciaosr_plan_init(ctypes.byref(d), _ptr(self.buf), nbytes.value,
                                            _stream(self.device)))
        self._ws = None

    # -- workspace (grown on demand, reused across calls on the same stream) ----
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def workspace_bytes(self, B, H, W, q, engine="auto"):
        n = ctypes.c_size_t(0)
        _lib.check(_lib.load().ciaosr_workspace_bytes(ctypes.byref(self.desc), B, H, W, q,
                                                      ENGINES[engine], ctypes.byref(n)))
        return n.value

    def engine_supported(self, engine):
        return _lib.load().ciaosr_engine_supported(ctypes.byref(self.desc), ENGINES[engine]) == 1

    # -- calls ---------------------------------------------------------------------
    def cross_scale_attention(self, feature, engine="auto"):
        """feature [B,C,H,W] -> [B,Cn,H,W] (CrossScaleAttention.forward)."""
        feature = _f32c(feature, "feature")
        B, C, H, W = feature.shape
        if C != self.channels:
            raise ValueError(f"feature has {C} channels, head was built for {self.channels}")
        out = torch.empty(B, self.n_nonlocal, H, W, dtype=torch.float32, device=feature.device)
        n = ctypes.c_size_t(0)
        _lib.check(_lib.load().ciaosr_cross_scale_attn_workspace_bytes(
            ctypes.byref(self.desc), B, H, W, ENGINES[engine], ctypes.byref(n)))
        ws = self._workspace(n.value)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ciaosr_cross_scale_attn_forward(
                ctypes.byref(self.desc), _ptr(self.buf), _ptr(feature), B, H, W, ENGINES[engine],
                _ptr(out), _ptr(ws), ws.numel(), _stream(self.device)))
        return out

    def query_rgb(self, feature, coord, cell, lr_image=None, nonlocal_feat=None, eval_bsize=None,
                  engine="auto"):
        """The head: feature [B,C,H,W], coord/cell [B,q,2] -> [B,q,3]."""
        feature = _f32c(feature, "feature")
        coord, cell = _f32c(coord, "coord"), _f32c(cell, "cell")
        B, C, H, W = feature.shape
        if C != self.channels:
            raise ValueError(f"feature has {C} channels, head was built for {self.channels}")
        if coord.dim() != 3 or coord.shape[-1] != 2 or coord.shape[0] != B:
            raise ValueError(f"coord must be [B,q,2] with B={B}, got {tuple(coord.shape)}")
        if cell.shape != coord.shape:
            raise ValueError(f"cell {tuple(cell.shape)} and coord {tuple(coord.shape)} differ")
        q = coord.shape[1]
        if lr_image is not None:
            lr_image = _f32c(lr_image, "lr_image")
            if tuple(lr_image.shape) != (B, 3, H, W):
                raise ValueError(f"lr_image must be [{B},3,{H},{W}], got {tuple(lr_image.shape)}")
        if nonlocal_feat is not None:
            nonlocal_feat = _f32c(nonlocal_feat, "nonlocal_feat")
            if tuple(nonlocal_feat.shape) != (B, self.n_nonlocal, H, W):
                raise ValueError("nonlocal_feat has the wrong shape")
        out = torch.empty(B, q, 3, dtype=torch.float32, device=feature.device)
        if q == 0:                       # an empty coordinate list gives an empty prediction, as in the reference
            return out
        nbytes = self.workspace_bytes(B, H, W, q, engine)
        ws = self._workspace(nbytes)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ciaosr_query_rgb_forward(
                ctypes.byref(self.desc), _ptr(self.buf), _ptr(feature), _ptr(nonlocal_feat),
                _ptr(coord), _ptr(cell), _ptr(lr_image), B, H, W, q,
                int(eval_bsize) if eval_bsize else 0, ENGINES[engine], _ptr(out), _ptr(ws),
                ws.numel(), _stream(self.device)))
        return out


def tile_blend_accumulate(tile_pred, acc, cnt, y0, x0, th, tw):
    """acc/cnt [B,3,Ho,Wo] += tile_pred [B, th*tw, 3] at (y0, x0) (ciaosr.py:253-255)."""
    tile_pred = _f32c(tile_pred, "tile_pred")
    B = tile_pred.shape[0]
    Ho, Wo = acc.shape[-2:]
    with torch.cuda.device(acc.device):
        _lib.check(_lib.load().ciaosr_tile_blend_accumulate(
            _ptr(tile_pred), B, th, tw, _ptr(acc), _ptr(cnt), Ho, Wo, y0, x0, _stream(acc.device)))


def tile_blend_finish(acc, cnt, mean=None, std=None, clamp01=False):
    """[B,3,Ho,Wo] accumulators -> [B, Ho*Wo, 3] = acc/cnt (* std + mean, clamped)."""
    B, _, Ho, Wo = acc.shape
    out = torch.empty(B, Ho * Wo, 3, dtype=torch.float32, device=acc.device)
    mean = _f32c(mean.reshape(3), "mean") if mean is not None else None
    std = _f32c(std.reshape(3), "std") if std is not None else None
    with torch.cuda.device(acc.device):
        _lib.check(_lib.load().ciaosr_tile_blend_finish(
            _ptr(acc), _ptr(cnt), B, Ho, Wo, _ptr(mean), _ptr(std), 1 if clamp01 else 0, _ptr(out),
            _stream(acc.device)))
    return out


class SplitTensor:
    """A [rows, features] fp32 matrix in the representation the tensor-core kernels compute in: fp16 `hi` and `lo`
    halves (x = hi + lo), row stride `ld` (multiple of 8), columns [features, ld) zero.  Produced by
    layernorm_split / window_attention(split_out=True) / LinearPlan.forward_split(split_out=True) and consumed by
    LinearPlan.forward_split, whose operand tiles are then TMA-loaded instead of re-split from fp32 per GEMM."""

    def __init__(self, rows, features, device):
        self.rows, self.features, self.ld = rows, features, (features + 7) // 8 * 8
        self.hi = torch.empty(rows, self.ld, dtype=torch.float16, device=device)
        self.lo = torch.empty(rows, self.ld, dtype=torch.float16, device=device)

    def float(self):
        return (self.hi.float() + self.lo.float())[:, :self.features]


def layernorm_split(x, ln):
    """nn.LayerNorm over the last dim -> SplitTensor [x.numel() / C, C] (ciaosr_layernorm_split_forward)."""
    x = _f32c(x, "x")
    c = x.shape[-1]
    out = SplitTensor(x.numel() // c, c, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ciaosr_layernorm_split_forward(
            _ptr(x), _ptr(_f32c(ln.weight.detach(), "weight")), _ptr(_f32c(ln.bias.detach(), "bias")), float(ln.eps),
            out.rows, c, _ptr(out.hi), _ptr(out.lo), out.ld, _stream(x.device)))
    return out


def window_attention_supported(c, heads, ws):
    threads = (ws * ws * heads + 31) // 32 * 32
    return heads > 0 and c % heads == 0 and c % 4 == 0 and c // heads <= 32 and ws * ws * heads <= 384 \
        and c // 2 <= threads


def window_attention(qkv, bias_table, h, w, heads, ws, shift, scale, split_out=False):
    """qkv [B, H*W, 3C] (natural token order) -> [B, H*W, C] (or a SplitTensor [B*H*W, C] with split_out): shifted-window
    multi-head attention with relative position bias and SW-MSA mask (ciaosr_window_attention_forward)."""
    qkv, bias_table = _f32c(qkv, "qkv"), _f32c(bias_table.detach(), "relative_position_bias_table")
    b, n, c3 = qkv.shape
    if n != h * w or c3 % 3:
        raise ValueError(f"qkv must be [B, {h * w}, 3C], got {tuple(qkv.shape)}")
    c = c3 // 3
    if tuple(bias_table.shape) != ((2 * ws - 1) ** 2, heads):
        raise ValueError(f"bias table must be [{(2 * ws - 1) ** 2}, {heads}], got {tuple(bias_table.shape)}")
    if split_out:
        out = SplitTensor(b * n, c, qkv.device)
        with torch.cuda.device(qkv.device):
            _lib.check(_lib.load().ciaosr_window_attention_split_forward(
                _ptr(qkv), _ptr(bias_table), b, h, w, c, heads, ws, shift, float(scale), _ptr(out.hi), _ptr(out.lo),
                out.ld, _stream(qkv.device)))
        return out
    out = torch.empty(b, n, c, dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().ciaosr_window_attention_forward(
            _ptr(qkv), _ptr(bias_table), b, h, w, c, heads, ws, shift, float(scale), _ptr(out), _stream(qkv.device)))
    return out


def layernorm(x, ln):
    """nn.LayerNorm over the last dim of a contiguous fp32 CUDA tensor (ciaosr_layernorm_forward)."""
    x = _f32c(x, "x")
    c = x.shape[-1]
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ciaosr_layernorm_forward(
            _ptr(x), _ptr(_f32c(ln.weight.detach(), "weight")), _ptr(_f32c(ln.bias.detach(), "bias")), float(ln.eps),
            x.numel() // c, c, _ptr(out), _stream(x.device)))
    return out


def profile_enable(on=True):
    """Bracket the library's stages with CUDA events (see ciaosr_profile_read)."""
    _lib.check(_lib.load().ciaosr_profile_enable(1 if on else 0))


def profile_read():
    """{stage: (milliseconds, launches)} accumulated since the last read; synchronises."""
    ms = (ctypes.c_float * _lib.N_STAGES)()
    n = (ctypes.c_int * _lib.N_STAGES)()
    _lib.check(_lib.load().ciaosr_profile_read(ms, n, _lib.N_STAGES))
    return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(_lib.STAGE_NAMES)}


def launch_count():
    return int(_lib.load().ciaosr_launch_count())


class CsAttnOnlyPlan(HeadPlan):
    """A plan for calling CrossScaleAttention as a standalone module.

    The C ABI describes a whole head; the three MLPs are filled with 1-wide
    zero layers that are never executed by ``cross_scale_attention``.
    """

    def __init__(self, params, channels, multi_scale=(2,), cs_softmax_scale=10.0):
        dev = params["cs_attn.down.weight"].device
        dk, cn = 9 * channels, channels * len(multi_scale)
        dv = dk + cn
        full = dict(params)

        def zeros(*shape):
            return torch.zeros(*shape, dtype=torch.float32, device=dev)

        for name, din, dout in (("imnet_k", dk + 4, dk), ("imnet_v", dv + 4, dv), ("imnet_q", dv, 3)):
            full[f"{name}.layers.0.weight"], full[f"{name}.layers.0.bias"] = zeros(1, din), zeros(1)
            full[f"{name}.layers.2.weight"], full[f"{name}.layers.2.bias"] = zeros(dout, 1), zeros(dout)
        super().__init__(full, channels, local_size=2, non_local_attn=True, multi_scale=multi_scale,
                         softmax_scale=1.0, cs_softmax_scale=cs_softmax_scale)


class RdnPlan:
    """Native RDN encoder (ciaosr_rdn_* in the C ABI): weights packed once, forward on the
    tensor cores with fp32-grade accuracy.  `params`: the generator's hoisted encoder
    state (sfe1.*, sfe2.*, rdbs.N.layers.M.conv.*, rdbs.N.lff.*, gff.0.*, gff.1.*)."""

    def __init__(self, params, mid_channels, channel_growth, num_blocks, num_layers):
        lib = _lib.load()
        self._keep = []
        self.device = params["sfe1.weight"].device
        if self.device.type != "cuda":
            raise RuntimeError("RdnPlan needs CUDA parameters")

        def p(key):
            t = _f32c(params[key].detach(), key)
            self._keep.append(t)
            return t.data_ptr()

        d = _lib.RdnDesc()
        d.abi_version = _lib.ABI_VERSION
        d.mid_channels, d.channel_growth = int(mid_channels), int(channel_growth)
        d.num_blocks, d.num_layers = int(num_blocks), int(num_layers)
        d.sfe1_w, d.sfe1_b, d.sfe2_w, d.sfe2_b = p("sfe1.weight"), p("sfe1.bias"), p("sfe2.weight"), p("sfe2.bias")
        n = num_blocks * num_layers
        self._dw = (ctypes.c_void_p * n)(*[p(f"rdbs.{r}.layers.{l}.conv.weight")
                                           for r in range(num_blocks) for l in range(num_layers)])
        self._db = (ctypes.c_void_p * n)(*[p(f"rdbs.{r}.layers.{l}.conv.bias")
                                           for r in range(num_blocks) for l in range(num_layers)])
        self._lw = (ctypes.c_void_p * num_blocks)(*[p(f"rdbs.{r}.lff.weight") for r in range(num_blocks)])
        self._lb = (ctypes.c_void_p * num_blocks)(*[p(f"rdbs.{r}.lff.bias") for r in range(num_blocks)])
        d.dense_w, d.dense_b = self._dw, self._db
        d.lff_w, d.lff_b = self._lw, self._lb
        d.gff0_w, d.gff0_b, d.gff1_w, d.gff1_b = p("gff.0.weight"), p("gff.0.bias"), p("gff.1.weight"), p("gff.1.bias")
        self.desc = d
        self.channels = int(mid_channels)
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.ciaosr_rdn_plan_bytes(ctypes.byref(d), ctypes.byref(nbytes)))
        with torch.cuda.device(self.device):
            self.buf = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=self.device)
            _lib.check(lib.ciaosr_rdn_plan_init(ctypes.byref(d), _ptr(self.buf), nbytes.value,
                                                _stream(self.device)))
        self._ws = None

    def forward(self, x):
        """x [B,3,H,W] -> feature [B,64,H,W]."""
        x = _f32c(x, "x")
        B, c, H, W = x.shape
        if c != 3:
            raise ValueError(f"RDN expects 3 input channels, got {c}")
        n = ctypes.c_size_t(0)
        _lib.check(_lib.load().ciaosr_rdn_workspace_bytes(ctypes.byref(self.desc), B, H, W, ctypes.byref(n)))
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = None
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, self.channels, H, W, dtype=torch.float32, device=x.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ciaosr_rdn_forward(ctypes.byref(self.desc), _ptr(self.buf), _ptr(x), B, H, W,
                                                      _ptr(out), _ptr(self._ws), self._ws.numel(),
                                                      _stream(self.device)))
        return out


class LinearPlan:
    """One nn.Linear packed for ``ciaosr_linear_forward`` (fp32-grade Linear on the tensor cores; the SwinIR
    trunk's qkv / proj / fc1 / fc2).  Holds references to the fp32 weight / bias it was packed from."""

    def __init__(self, weight, bias):
        lib = _lib.load()
        self.weight = _f32c(weight.detach(), "weight")
        # a private copy of the bias: the epilogue reads it as float4, so it must be 16-byte aligned
        self.bias = _f32c(bias.detach(), "bias").clone() if bias is not None else None
        self.device = self.weight.device
        d = _lib.LinearDesc()
        d.abi_version = _lib.ABI_VERSION
        d.out_features, d.in_features = self.weight.shape
        d.weight = self.weight.data_ptr()
        d.bias = self.bias.data_ptr() if self.bias is not None else None
        self.desc = d
        n = ctypes.c_size_t(0)
        _lib.check(lib.ciaosr_linear_plan_bytes(ctypes.byref(d), ctypes.byref(n)))
        with torch.cuda.device(self.device):
            self.buf = torch.empty(max(n.value, 256), dtype=torch.uint8, device=self.device)
            _lib.check(lib.ciaosr_linear_plan_init(ctypes.byref(d), _ptr(self.buf), n.value, _stream(self.device)))

    @staticmethod
    def supports(weight):
        return weight.is_cuda and weight.dtype == torch.float32 and weight.shape[0] % 4 == 0 and weight.shape[1] % 4 == 0

    def forward(self, x, gelu=False, residual=None):
        """x [..., in_features] fp32 CUDA -> [..., out_features] (+ residual of that shape, added in the epilogue)."""
        x = _f32c(x, "x")
        k, n = self.desc.in_features, self.desc.out_features
        if x.shape[-1] != k:
            raise ValueError(f"linear expects {k} input features, got {x.shape[-1]}")
        rows = x.numel() // k
        out = torch.empty(*x.shape[:-1], n, dtype=torch.float32, device=x.device)
        if residual is not None:
            residual = _f32c(residual, "residual")
            if residual.shape != out.shape:
                raise ValueError(f"residual {tuple(residual.shape)} does not match the output {tuple(out.shape)}")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ciaosr_linear_forward_res(ctypes.byref(self.desc), _ptr(self.buf), _ptr(x), rows,
                                                             1 if gelu else 0, _ptr(residual), _ptr(out),
                                                             _stream(self.device)))
        return out


def linear_plan_for(lin):
    """The cached LinearPlan of an nn.Linear (rebuilt when its parameters change or move), or None if unsupported."""
    w = lin.weight
    if not LinearPlan.supports(w):
        return None
    key = (w.data_ptr(), w._version, None if lin.bias is None else (lin.bias.data_ptr(), lin.bias._version))
    mc = module_cache(lin)
    cache = mc.get("linear_plan")
    if cache is None or cache[0] != key:
        cache = mc["linear_plan"] = (key, LinearPlan(w, lin.bias))
    return cache[1]


def conv3x3_plan_for(conv):
    """The cached Conv3x3Plan of an nn.Conv2d (rebuilt when its parameters change or move), or None when the native
    path does not support the layer."""
    if not Conv3x3Plan.supports(conv):
        return None
    w = conv.weight
    key = (w.data_ptr(), w._version, None if conv.bias is None else (conv.bias.data_ptr(), conv.bias._version))
    mc = module_cache(conv)
    cache = mc.get("conv_plan")
    if cache is None or cache[0] != key:
        cache = mc["conv_plan"] = (key, Conv3x3Plan(w, conv.bias))
    return cache[1]


def _linear_forward_split(self, a, gelu=False, residual=None, split_out=False, out_shape=None):
    """`a` SplitTensor [rows, in_features] -> fp32 [*out_shape or (rows,), out_features] (+ residual), or a SplitTensor
    with split_out (ciaosr_linear_forward_split: TMA-fed operand, nothing re-split in the GEMM)."""
    k, n = self.desc.in_features, self.desc.out_features
    if a.features != k:
        raise ValueError(f"linear expects {k} input features, got {a.features}")
    out = out_hi = out_lo = None
    ldo = 0
    if split_out:
        res = SplitTensor(a.rows, n, a.hi.device)
        out_hi, out_lo, ldo = res.hi, res.lo, res.ld
    else:
        res = out = torch.empty(*(out_shape or (a.rows,)), n, dtype=torch.float32, device=a.hi.device)
    if residual is not None:
        residual = _f32c(residual, "residual")
        if residual.numel() != a.rows * n:
            raise ValueError("residual does not match the output")
    with torch.cuda.device(self.device):
        _lib.check(_lib.load().ciaosr_linear_forward_split(
            ctypes.byref(self.desc), _ptr(self.buf), _ptr(a.hi), _ptr(a.lo), a.ld, a.rows, 1 if gelu else 0,
            _ptr(residual), _ptr(out), _ptr(out_hi), _ptr(out_lo), ldo, _stream(self.device)))
    return res


LinearPlan.forward_split = _linear_forward_split


class Conv3x3Plan:
    """One nn.Conv2d(Cin, Cout, 3, 1, 1) packed for ``ciaosr_conv3x3_nhwc_forward`` (implicit GEMM on the tensor
    cores over NHWC maps = token tensors; the SwinIR trunk's RSTB / after-body convolutions)."""

    def __init__(self, weight, bias):
        lib = _lib.load()
        self.weight = _f32c(weight.detach(), "weight")
        self.bias = _f32c(bias.detach(), "bias").clone() if bias is not None else None     # float4-aligned copy
        self.device = self.weight.device
        d = _lib.Conv3x3Desc()
        d.abi_version = _lib.ABI_VERSION
        d.out_channels, d.in_channels = self.weight.shape[:2]
        d.weight = self.weight.data_ptr()
        d.bias = self.bias.data_ptr() if self.bias is not None else None
        self.desc = d
        n = ctypes.c_size_t(0)
        _lib.check(lib.ciaosr_conv3x3_plan_bytes(ctypes.byref(d), ctypes.byref(n)))
        with torch.cuda.device(self.device):
            self.buf = torch.empty(max(n.value, 256), dtype=torch.uint8, device=self.device)
            _lib.check(lib.ciaosr_conv3x3_plan_init(ctypes.byref(d), _ptr(self.buf), n.value, _stream(self.device)))

    @staticmethod
    def supports(conv):
        w = conv.weight
        return (isinstance(conv, torch.nn.Conv2d) and w.is_cuda and w.dtype == torch.float32
                and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (1, 1)
                and tuple(conv.dilation) == (1, 1) and conv.groups == 1 and conv.padding_mode == "zeros"
                and w.shape[0] % 4 == 0 and w.shape[1] % 4 == 0)

    def forward(self, x, residual=None, relu=False):
        """x [B, H, W, Cin] (NHWC) -> [B, H, W, Cout]: act(conv(x) + bias) (+ residual)."""
        x = _f32c(x, "x")
        b, h, w, c = x.shape
        if c != self.desc.in_channels:
            raise ValueError(f"conv expects {self.desc.in_channels} input channels, got {c}")
        out = torch.empty(b, h, w, self.desc.out_channels, dtype=torch.float32, device=x.device)
        if residual is not None:
            residual = _f32c(residual, "residual")
            if residual.numel() != out.numel():
                raise ValueError("residual does not match the output")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ciaosr_conv3x3_nhwc_forward(ctypes.byref(self.desc), _ptr(self.buf), _ptr(x), b, h,
                                                               w, 2 if relu else 0, _ptr(residual), _ptr(out),
                                                               _stream(self.device)))
        return out
