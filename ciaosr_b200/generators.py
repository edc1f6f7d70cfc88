"""The drop-in boundary: the reference's generator classes over the native head.

Mirrors ``LocalImplicitSRNet`` and its RDN / EDSR / SwinIR subclasses
(mmedited/models/backbones/sr_backbones/ciaosr_net.py:17-525): same constructor
keywords (plus the six keywords the 002 configs pass),
same ``state_dict`` keys, same ``forward(x, coord, cell, test_mode)`` ->
``[B, Q, 3]``, same ``init_weights(pretrained, strict)`` error behaviour.

What differs is where the work happens: ``gen_feature`` (the encoder) stays in
PyTorch; everything after it -- unfold, cross-scale attention, local-ensemble
encoding, the three MLPs, inner attention, bilinear residual -- is one call
into ``libciaosr_b200.so``.  There is no PyTorch/CPU fallback for that part.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native
from .builder import build_backbone, build_component, load_checkpoint
from .cross_scale_attention import CrossScaleAttention


class LocalImplicitSRNet(nn.Module):
    """Base generator (ciaosr_net.py:17-264).

    Args follow the reference; ``engine`` ('auto' | 'simt' | 'tcgen05') selects the
    native engine and is the only addition.
    """

    def __init__(self, encoder, imnet_q, imnet_k, imnet_v, query_mlp=None, key_mlp=None,
                 value_mlp=None, local_size=2, feat_unfold=True, eval_bsize=None,
                 non_local_attn=True, multi_scale=[2], softmax_scale=1,
                 # keywords the 002 configs pass (configs/002_real_wogan...py:54-59)
                 local_ensemble_coord=True, imnet_k_type=None, imnet_v_type=None, res=True,
                 cat_nla_v=None, engine="auto", cuda_graph=False, channels_last=True):
        super().__init__()
        self.feat_unfold = feat_unfold
        self.eval_bsize = eval_bsize
        self.local_size = local_size
        self.non_local_attn = non_local_attn
        self.multi_scale = list(multi_scale)
        self.softmax_scale = softmax_scale
        self.res = res
        self.engine = engine
        # plumbing knobs (no effect on results): replay the whole inference forward as one CUDA graph per
        # input shape (the PyTorch encoder is ~450 small launches and is CPU-launch-bound otherwise), and
        # run the encoder convolutions in channels_last (cuDNN's native layout; saves the NCHW<->NHWC passes)
        self.cuda_graph = cuda_graph
        self.channels_last = channels_last
        if not local_ensemble_coord:
            raise NotImplementedError("local_ensemble_coord=False has no counterpart in the reference head")

        self.encoder = build_backbone(encoder)
        imnet_dim = self.encoder.mid_channels if hasattr(self.encoder, "mid_channels") \
            else self.encoder.embed_dim
        self.imnet_dim = imnet_dim
        # the reference mutates the config dicts in place (ciaosr_net.py:61-76); so do we
        unit = imnet_dim * 9 if feat_unfold else imnet_dim
        imnet_q["in_dim"] = unit
        imnet_k["in_dim"] = imnet_k["out_dim"] = unit
        imnet_v["in_dim"] = imnet_v["out_dim"] = unit
        imnet_k["in_dim"] += 4
        imnet_v["in_dim"] += 4
        if non_local_attn:
            extra = imnet_dim * len(self.multi_scale)
            imnet_q["in_dim"] += extra
            imnet_v["in_dim"] += extra
            imnet_v["out_dim"] += extra
        self.imnet_q = build_component(imnet_q)
        self.imnet_k = build_component(imnet_k)
        self.imnet_v = build_component(imnet_v)
        if non_local_attn:
            self.cs_attn = CrossScaleAttention(channel=imnet_dim, scale=self.multi_scale)

    # -- native state ----------------------------------------------------------------
    # Packed plans and CUDA graphs are kept outside the module (native.module_cache): they hold ctypes structs and
    # device pointers, which must not travel through copy.deepcopy / pickle (EMA copies, torch.save(model)).
    def _nc(self):
        return native.module_cache(self)

    @property
    def _plan(self):
        return self._nc().get("plan")

    @property
    def _graphs(self):
        return self._nc().setdefault("graphs", {})

    def _native_buffers(self):
        """Every device buffer the library was handed on behalf of this module right now: packed plans and their
        current workspaces.  A captured CUDA graph addresses them by raw pointer, so each graph entry holds these
        references: a later, larger call may replace a plan's workspace (or a parameter update its packed buffer),
        but the blocks an existing graph replays into stay allocated for as long as that graph does."""
        keep = []
        for m in self.modules():
            for v in native.module_cache(m).values():
                plan = v[1] if isinstance(v, tuple) and len(v) == 2 else v
                if isinstance(plan, (native.HeadPlan, native.RdnPlan, native.LinearPlan, native.Conv3x3Plan)):
                    keep += [plan, plan.buf, getattr(plan, "_ws", None)]
        return keep

    # -- native plan -----------------------------------------------------------------
    def _head_params(self):
        prefixes = ("imnet_q.", "imnet_k.", "imnet_v.", "cs_attn.")
        return {k: v for k, v in self.state_dict().items() if k.startswith(prefixes)}

    def head_plan(self):
        """Packed weights for the kernels; rebuilt when a parameter changes or moves."""
        params = self._head_params()
        key = tuple((k, v.data_ptr(), v._version, str(v.device)) for k, v in params.items())
        nc = self._nc()
        if nc.get("plan") is None or key != nc.get("plan_key"):
            nc["plan"] = native.HeadPlan(
                params, self.imnet_dim, local_size=self.local_size,
                non_local_attn=self.non_local_attn, multi_scale=self.multi_scale,
                softmax_scale=float(self.softmax_scale), feat_unfold=self.feat_unfold,
                cs_softmax_scale=float(self.cs_attn.softmax_scale) if self.non_local_attn else 10.0)
            nc["plan_key"] = key
        return nc["plan"]

    # -- reference API -----------------------------------------------------------------
    def forward(self, x, coord, cell, test_mode=False):
        """x [B,3,H,W] normalised LR, coord/cell [B,Q,2] (y,x) -> [B,Q,3]."""
        if torch.is_grad_enabled() and not test_mode and (x.requires_grad or any(
                p.requires_grad for p in self.parameters())):
            return self._forward_train(x, coord, cell)
        with torch.no_grad():
            if self.cuda_graph and x.is_cuda:
                return self._forward_graphed(x, coord, cell, test_mode)
            return self._forward_eager(x, coord, cell, test_mode)

    def _forward_train(self, x, coord, cell):
        """Training forward (ciaosr.py:88 -> ciaosr_net.py:98-108 with test_mode=False: no eval_bsize chunking): the
        encoder runs in PyTorch under autograd, the head's VALUE comes from the native kernels and its gradient from
        a recompute with differentiable torch ops (head_autograd.HeadFunction)."""
        from .head_autograd import HeadFunction
        if not x.is_cuda:
            raise RuntimeError("ciaosr_b200 has no CPU path for the head (training forward included)")
        feature = self.gen_feature(x)
        if len(feature) != 1:
            raise NotImplementedError("every reference encoder returns exactly one feature map")
        params = self._head_params_live()
        hyper = dict(engine=self.engine, local_size=self.local_size, non_local_attn=self.non_local_attn,
                     softmax_scale=float(self.softmax_scale),
                     cs_softmax_scale=float(self.cs_attn.softmax_scale) if self.non_local_attn else 10.0)
        return HeadFunction.apply(self.head_plan(), hyper, feature[0].contiguous(), coord.contiguous(),
                                  cell.contiguous(), x.detach().contiguous() if self.res else None,
                                  tuple(params), *params.values())

    def _head_params_live(self):
        """The head's parameters / buffers themselves (not detached copies), keyed like the state_dict."""
        prefixes = ("imnet_q.", "imnet_k.", "imnet_v.", "cs_attn.")
        live = dict(self.named_parameters())
        live.update(dict(self.named_buffers()))
        return {k: v for k, v in live.items() if k.startswith(prefixes)}

    def _forward_eager(self, x, coord, cell, test_mode):
        xin = x.contiguous(memory_format=torch.channels_last) if (self.channels_last and x.is_cuda) else x
        feature = [f.contiguous() for f in self.gen_feature(xin)]
        chunk = self.eval_bsize if (self.eval_bsize is not None and test_mode) else None
        return self.query_rgb(feature, coord, cell, lr_image=x.contiguous() if self.res else None,
                              eval_bsize=chunk)

    def _forward_graphed(self, x, coord, cell, test_mode):
        """One CUDA graph per (shapes, test_mode); inputs are copied into static buffers and the
        static output is cloned, so callers see ordinary tensors."""
        # Head AND encoder weights are packed into plans (HeadPlan, RdnPlan, one LinearPlan per trunk Linear) that the
        # captured kernels address directly: a new graph is needed when ANY parameter or buffer changes or moves
        # (load_state_dict, init_weights, an in-place update), not only the head's.
        sig = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        key = (tuple(x.shape), tuple(coord.shape), bool(test_mode), x.device.index, sig)
        entry = self._graphs.get(key)
        if entry is None:
            sx, sc, sl = x.clone(), coord.clone(), cell.clone()
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):                      # warm-up: plan, workspace, cuDNN algos
                for _ in range(2):
                    self._forward_eager(sx, sc, sl, test_mode)
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                so = self._forward_eager(sx, sc, sl, test_mode)
            # the plans / workspaces this capture recorded pointers of stay alive with the graph (_native_buffers)
            entry = (graph, sx, sc, sl, so, self._native_buffers())
            if len(self._graphs) >= 8:                         # bound the pools kept alive
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = entry
        graph, sx, sc, sl, so = entry[:5]
        sx.copy_(x, non_blocking=True)
        sc.copy_(coord, non_blocking=True)
        sl.copy_(cell, non_blocking=True)
        graph.replay()
        return so.clone()

    def query_rgb(self, features, coord, scale=None, lr_image=None, eval_bsize=None):
        """ciaosr_net.py:113-224 (`scale` is the reference's name for the cell tensor)."""
        if not isinstance(features, (list, tuple)):
            features = [features]
        if len(features) != 1:
            raise NotImplementedError("every reference encoder returns exactly one feature map")
        return self.head_plan().query_rgb(features[0], coord, scale, lr_image=lr_image,
                                          eval_bsize=eval_bsize, engine=self.engine)

    def batched_predict(self, x, coord, cell):
        """ciaosr_net.py:226-248.  The native call handles the whole query axis at once;
        `eval_bsize` only selects which cell seeds tx/ty, exactly like the chunk loop."""
        return self.query_rgb(x, coord, cell, eval_bsize=self.eval_bsize)

    def init_weights(self, pretrained=None, strict=True):
        if isinstance(pretrained, str):
            load_checkpoint(self, pretrained, strict=strict)
        elif pretrained is not None:
            raise TypeError('"pretrained" must be a str or None. '
                            f'But received {type(pretrained)}.')

    def gen_feature(self, x):
        raise NotImplementedError


class LocalImplicitSRRDN(LocalImplicitSRNet):
    """ciaosr_net.py:267-342."""

    def __init__(self, encoder, imnet_q, imnet_k, imnet_v, **kwargs):
        super().__init__(encoder=encoder, imnet_q=imnet_q, imnet_k=imnet_k, imnet_v=imnet_v, **kwargs)
        self.sfe1 = self.encoder.sfe1
        self.sfe2 = self.encoder.sfe2
        self.rdbs = self.encoder.rdbs
        self.gff = self.encoder.gff
        self.num_blocks = self.encoder.num_blocks
        self._enc_geom = (self.encoder.mid_channels, self.encoder.channel_growth, self.encoder.num_blocks,
                          self.encoder.num_layers)
        del self.encoder
        # encoder fast path: the same RDN on the tensor cores with fp32-grade accuracy.
        # 'auto' uses it for CUDA inputs when the geometry admits it (mid_channels == growth == 64);
        # False keeps the PyTorch / cuDNN encoder of the reference.
        self.native_encoder = "auto"

    def _native_encoder_plan(self):
        names = ("sfe1.", "sfe2.", "rdbs.", "gff.")
        params = {k: v for k, v in self.state_dict().items() if k.startswith(names)}
        key = tuple((k, v.data_ptr(), v._version) for k, v in params.items())
        nc = self._nc()
        if nc.get("enc_plan") is None or key != nc.get("enc_key"):
            mid, growth, nb, nl = self._enc_geom
            nc["enc_plan"] = native.RdnPlan(params, mid, growth, nb, nl)
            nc["enc_key"] = key
        return nc["enc_plan"]

    def gen_feature(self, x):
        mid, growth = self._enc_geom[:2]
        if self.native_encoder and x.is_cuda and mid == 64 and growth == 64 and not torch.is_grad_enabled():
            return [self._native_encoder_plan().forward(x)]
        if self.native_encoder is True:
            raise RuntimeError("native_encoder=True needs a CUDA input, no_grad, and mid_channels == growth == 64")
        sfe1 = self.sfe1(x)
        x = self.sfe2(sfe1)
        local_features = []
        for i in range(self.num_blocks):
            x = self.rdbs[i](x)
            local_features.append(x)
        return [self.gff(torch.cat(local_features, 1)) + sfe1]


class LocalImplicitSREDSR(LocalImplicitSRNet):
    """ciaosr_net.py:345-408."""

    def __init__(self, encoder, imnet_q, imnet_k, imnet_v, **kwargs):
        super().__init__(encoder=encoder, imnet_q=imnet_q, imnet_k=imnet_k, imnet_v=imnet_v, **kwargs)
        self.conv_first = self.encoder.conv_first
        self.body = self.encoder.body
        self.conv_after_body = self.encoder.conv_after_body
        del self.encoder
        # encoder fast path: the 64 -> 64 convolutions of the residual trunk as implicit GEMMs on the
        # tensor cores with fp32-grade accuracy (csrc/linear_tc.cu: ciaosr_conv3x3_nhwc_forward, ReLU / residual fused).
        # 'auto' = on for CUDA inference; False keeps the PyTorch / cuDNN trunk of the reference.
        self.native_encoder = "auto"

    def _gen_feature_native(self, x):
        """ciaosr_net.py:393-408 on NHWC maps: conv_first stays in PyTorch (3 input channels; strict fp32), every
        other convolution runs natively; None when a layer is outside what the native path supports."""
        plans = []
        for blk in self.body:
            p1, p2 = native.conv3x3_plan_for(blk.conv1), native.conv3x3_plan_for(blk.conv2)
            if p1 is None or p2 is None or float(getattr(blk, "res_scale", 1.0)) != 1.0:
                return None
            plans.append((p1, p2))
        last = native.conv3x3_plan_for(self.conv_after_body)
        if last is None:
            return None
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            first = self.conv_first(x).permute(0, 2, 3, 1).contiguous()               # [B, H, W, C]
        t = first
        for p1, p2 in plans:
            t = p2.forward(p1.forward(t, relu=True), residual=t)                      # x + conv2(relu(conv1(x)))
        return last.forward(t, residual=first).permute(0, 3, 1, 2)

    def gen_feature(self, x):
        if self.native_encoder and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled():
            res = self._gen_feature_native(x)
            if res is not None:
                return [res]
        if self.native_encoder is True:
            raise RuntimeError("native_encoder=True needs a CUDA fp32 input, no_grad, and 3x3 / stride-1 / res_scale-1 "
                               "convolutions with channel counts that are multiples of 4")
        x = self.conv_first(x)
        res = self.conv_after_body(self.body(x))
        res += x
        return [res]


class LocalImplicitSRSWINIR(LocalImplicitSRNet):
    """ciaosr_net.py:411-525."""

    def __init__(self, window_size, encoder, imnet_q, imnet_k, imnet_v, **kwargs):
        super().__init__(encoder=encoder, imnet_q=imnet_q, imnet_k=imnet_k, imnet_v=imnet_v, **kwargs)
        self.window_size = window_size
        self.conv_first = self.encoder.conv_first
        self.patch_embed = self.encoder.patch_embed
        self.pos_drop = self.encoder.pos_drop
        self.layers = self.encoder.layers
        self.norm = self.encoder.norm
        self.patch_unembed = self.encoder.patch_unembed
        self.conv_after_body = self.encoder.conv_after_body
        del self.encoder
        # encoder fast path: the trunk's Linear layers through the library's fp32-grade
        # tensor-core Linear (swinir.native_linear).  'auto' = on for CUDA inference; False keeps plain PyTorch.
        self.native_encoder = "auto"

    def forward_features(self, x):
        x_size = (x.shape[2], x.shape[3])
        x = self.pos_drop(self.patch_embed(x))
        for layer in self.layers:
            x = layer(x, x_size)
        from . import swinir
        return self.patch_unembed(swinir._norm(x, self.norm), x_size)

    def gen_feature(self, img):
        _, _, h, w = img.size()
        pad_h = (self.window_size - h % self.window_size) % self.window_size
        pad_w = (self.window_size - w % self.window_size) % self.window_size
        from . import swinir
        x = self.conv_first(F.pad(img, (0, pad_w, 0, pad_h), "reflect"))
        with swinir.native_linear(bool(self.native_encoder) and img.is_cuda and not torch.is_grad_enabled()):
            res = self._after_body_native(x)
            if res is None:
                res = self.conv_after_body(self.forward_features(x))
                res += x
        return [res[:, :, :h, :w]]

    def _after_body_native(self, x):
        """conv_after_body(forward_features(x)) + x with the 3x3 convolution on the token tensor (an NHWC map) through
        the native implicit-GEMM path; None when that path does not apply (then the PyTorch ops above run)."""
        from . import swinir
        if not (swinir._NATIVE and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled()
                and isinstance(self.conv_after_body, nn.Conv2d) and native.Conv3x3Plan.supports(self.conv_after_body)):
            return None
        x_size = (x.shape[2], x.shape[3])
        t = self.pos_drop(self.patch_embed(x))
        for layer in self.layers:
            t = layer(t, x_size)
        y = swinir._conv3x3_tokens(swinir._norm(t, self.norm), x_size, self.conv_after_body,
                                   residual=x.permute(0, 2, 3, 1))
        if y is None:
            return None
        return y.view(x.shape[0], x_size[0], x_size[1], -1).permute(0, 3, 1, 2)
