"""CrossScaleAttention: the reference's module interface
(mmedited/models/common/arch_csnln.py:407-532) over the native kernels."""
import torch
import torch.nn as nn

from . import native


class _ConvPReLU(nn.Sequential):
    """BasicBlock(conv 1x1, PReLU) -> state_dict keys ``0.weight, 0.bias, 1.weight``."""

    def __init__(self, cin, cout):
        super().__init__(nn.Conv2d(cin, cout, 1, padding=0, bias=True), nn.PReLU())


class CrossScaleAttention(nn.Module):
    def __init__(self, channel=64, reduction=2, ksize=3, scale=2, stride=1, softmax_scale=10,
                 average=True):
        super().__init__()
        if ksize != 3 or stride != 1 or not average or reduction != 2:
            raise NotImplementedError("only ksize=3, stride=1, reduction=2, average=True (the reference's use)")
        self.ksize, self.stride, self.softmax_scale, self.average = ksize, stride, softmax_scale, average
        self.scale = list(scale) if isinstance(scale, (list, tuple)) else [scale]
        self.channel = channel
        self.register_buffer("escape_NaN", torch.FloatTensor([1e-4]))
        self.conv_match_1 = _ConvPReLU(channel, channel // reduction)
        self.conv_match_2 = _ConvPReLU(channel, channel // reduction)
        self.conv_assembly = _ConvPReLU(channel, channel)
        if 3 in self.scale:
            self.downx3 = nn.Conv2d(channel, channel, ksize, 3, 1)
        if 4 in self.scale:
            self.downx4 = nn.Conv2d(channel, channel, ksize, 4, 1)
        self.down = nn.Conv2d(channel, channel, ksize, 2, 1)

    @property
    def _plan(self):
        """(parameter versions, CsAttnOnlyPlan) of the standalone path; held outside the module (native.module_cache)."""
        return native.module_cache(self).get("plan")

    @_plan.setter
    def _plan(self, value):
        native.module_cache(self)["plan"] = value

    def forward(self, x):
        """[B,C,H,W] -> [B, C*len(scale), H, W]; standalone use (inside the head the
        generator's own plan computes it)."""
        if self._plan is None or self._plan[0] != _param_versions(self):
            params = {f"cs_attn.{k}": v for k, v in self.state_dict().items()}
            # a head plan needs the three MLPs; build a minimal one around this module
            self._plan = (_param_versions(self), native.CsAttnOnlyPlan(params, self.channel, self.scale,
                                                                      self.softmax_scale))
        return self._plan[1].cross_scale_attention(x)


def _param_versions(module):
    return tuple((p.data_ptr(), p._version) for p in list(module.parameters()) + list(module.buffers()))
