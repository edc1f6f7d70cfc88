"""Import harness for the UNMODIFIED reference (test infrastructure only).

This file is not product code.  It exists so that, in the build container
where ``/root/reference`` is mounted read-only, the reference's own PyTorch
implementation of the hot path can be executed to (a) mint golden vectors
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``) and (b) pin the CPU
restatement in ``oracle/ciaosr_oracle.py``.

The reference depends on mmcv-full / mmedit 0.11 / thop / timm, none of which
is installed and none of which can be installed (no network).  Only a handful
of names are touched on the hot path, so they are provided as stub modules in
``sys.modules`` before the reference is imported.  Nothing is copied from the
reference: it is imported from where it lies.

Stubbed names and the reference lines that import them:
  mmcv.runner.load_checkpoint, mmcv.cnn.constant_init   ciaosr_net.py:4-5
  mmcv.runner.auto_fp16                                 basic_restorer.py:7
  mmedit.utils.get_root_logger                          ciaosr_net.py:6
  mmedit.datasets.pipelines.utils.make_coord            ciaosr_net.py:7, ciaosr.py:11
  mmedit.models.builder.build_backbone/build_component  ciaosr_net.py:8
  mmedit.models.builder.build_loss, mmedit.models.base.BaseModel
                                                        basic_restorer.py:10-11
  mmedit.core.{tensor2img, psnr, ssim}                  basic_restorer.py:9
  thop.profile                                          ciaosr.py:15
  timm.models.layers.{DropPath, to_2tuple, trunc_normal_}  swinir_net.py:11
The reference's SwinIR also calls ``.cuda()`` on sub-modules inside its constructor
(swinir_net.py:684,723,725); ``build_reference_swinir_generator`` neutralises ``nn.Module.cuda``
for the duration of that constructor so the trunk can be built in the GPU-less container.

``/root/reference`` does not exist on the GPU box; everything that imports
this module must be guarded by :func:`reference_available`.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    """Where the unmodified reference package lies: the mounted checkout in the build container, else the
    runtime copy `__graft_entry__.build()` places under baseline/_ref (git-ignored; it travels to the GPU box so
    that `bench.py --impl reference` can time the reference itself there)."""
    if os.environ.get("CIAOSR_NO_REFERENCE"):              # tests: exercise the "no copy of the reference" path
        return os.environ.get("CIAOSR_REFERENCE_ROOT", "/nonexistent")
    for cand in (os.environ.get("CIAOSR_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "mmedited")):
            return cand
    return os.environ.get("CIAOSR_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _reference_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mmedited"))


# ----------------------------------------------------------------------------
# third-party pieces restated (mmedit 0.11 is not vendored by the reference)
# ----------------------------------------------------------------------------
def make_coord(shape, ranges=None, flatten=True):
    """mmedit.datasets.pipelines.utils.make_coord.

    Pixel-centre coordinates in [-1, 1] per dim, (y, x) order, fp32.
    """
    seqs = []
    for i, n in enumerate(shape):
        v0, v1 = (-1, 1) if ranges is None else ranges[i]
        r = (v1 - v0) / (2 * n)
        seqs.append(v0 + r + (2 * r) * torch.arange(n).float())
    ret = torch.stack(torch.meshgrid(*seqs, indexing="ij"), dim=-1)
    if flatten:
        ret = ret.view(-1, ret.shape[-1])
    return ret


class _ResBlockNoBN(nn.Module):
    def __init__(self, c, res_scale=1.0):
        super().__init__()
        self.res_scale = res_scale
        self.conv1 = nn.Conv2d(c, c, 3, 1, 1)
        self.conv2 = nn.Conv2d(c, c, 3, 1, 1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return x + self.conv2(self.relu(self.conv1(x))) * self.res_scale


class StubEDSR(nn.Module):
    """Attribute-compatible EDSR trunk (ciaosr_net.py:388-391 hoists these)."""

    def __init__(self, in_channels=3, out_channels=3, mid_channels=64,
                 num_blocks=16, **_unused):
        super().__init__()
        self.mid_channels = mid_channels
        self.conv_first = nn.Conv2d(in_channels, mid_channels, 3, 1, 1)
        self.body = nn.Sequential(*[_ResBlockNoBN(mid_channels)
                                    for _ in range(num_blocks)])
        self.conv_after_body = nn.Conv2d(mid_channels, mid_channels, 3, 1, 1)


class _DenseLayer(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return torch.cat([x, self.relu(self.conv(x))], 1)


class _RDB(nn.Module):
    def __init__(self, cin, growth, num_layers):
        super().__init__()
        self.layers = nn.Sequential(*[_DenseLayer(cin + growth * i, growth)
                                      for i in range(num_layers)])
        self.lff = nn.Conv2d(cin + growth * num_layers, cin, 1)

    def forward(self, x):
        return x + self.lff(self.layers(x))


class StubRDN(nn.Module):
    """Attribute-compatible RDN trunk (ciaosr_net.py:314-318 hoists these)."""

    def __init__(self, in_channels=3, out_channels=3, mid_channels=64,
                 num_blocks=16, upscale_factor=4, num_layers=8,
                 channel_growth=64, **_unused):
        super().__init__()
        self.mid_channels = mid_channels
        self.num_blocks = num_blocks
        self.sfe1 = nn.Conv2d(in_channels, mid_channels, 3, padding=1)
        self.sfe2 = nn.Conv2d(mid_channels, mid_channels, 3, padding=1)
        self.rdbs = nn.ModuleList([_RDB(mid_channels, channel_growth, num_layers)
                                   for _ in range(num_blocks)])
        self.gff = nn.Sequential(
            nn.Conv2d(mid_channels * num_blocks, mid_channels, 1),
            nn.Conv2d(mid_channels, mid_channels, 3, padding=1))


_STRING_TYPES = {"EDSR": StubEDSR, "RDN": StubRDN}


def _build(cfg):
    cfg = dict(cfg)
    typ = cfg.pop("type")
    if isinstance(typ, str):
        if typ == "MLPRefiner":
            typ = import_reference().mlp.MLPRefiner
        elif typ == "L1Loss":
            return nn.L1Loss()
        else:
            typ = _STRING_TYPES[typ]
    return typ(**cfg)


def _install_stubs():
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_ciaosr_stub", False):
        return

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    mmcv = mod("mmcv")
    mmcv._ciaosr_stub = True
    mmcv.imwrite = lambda *a, **k: None
    runner = mod("mmcv.runner")
    runner.load_checkpoint = lambda *a, **k: None
    runner.auto_fp16 = lambda *a, **k: (lambda f: f)
    cnn = mod("mmcv.cnn")
    cnn.constant_init = lambda *a, **k: None
    mmcv.runner, mmcv.cnn = runner, cnn

    mmedit = mod("mmedit")
    utils = mod("mmedit.utils")
    utils.get_root_logger = lambda *a, **k: None
    datasets = mod("mmedit.datasets")
    pipelines = mod("mmedit.datasets.pipelines")
    putils = mod("mmedit.datasets.pipelines.utils")
    putils.make_coord = make_coord
    models = mod("mmedit.models")
    builder = mod("mmedit.models.builder")
    builder.build_backbone = _build
    builder.build_component = _build
    builder.build_loss = _build
    base = mod("mmedit.models.base")
    base.BaseModel = nn.Module
    core = mod("mmedit.core")
    core.tensor2img = lambda *a, **k: None
    core.psnr = lambda *a, **k: 0.0
    core.ssim = lambda *a, **k: 0.0
    mmedit.utils, mmedit.datasets, mmedit.models, mmedit.core = utils, datasets, models, core
    datasets.pipelines = pipelines
    pipelines.utils = putils
    models.builder, models.base = builder, base

    thop = mod("thop")
    thop.profile = lambda *a, **k: (0, 0)

    class _DropPath(nn.Module):                      # inference: identity
        def __init__(self, drop_prob=None):
            super().__init__()

        def forward(self, x):
            return x

    timm = mod("timm")
    tmodels = mod("timm.models")
    tlayers = mod("timm.models.layers")
    tlayers.DropPath = _DropPath
    tlayers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    tlayers.trunc_normal_ = nn.init.trunc_normal_
    timm.models, tmodels.layers = tmodels, tlayers


_REF = None


def import_reference():
    """Return the reference modules on the hot path (imported in place).

    This repo ships a compat namespace package with the same name (``mmedited``);
    the reference's copy is imported with the reference root first on ``sys.path``
    and then removed from ``sys.modules`` again, so both can live in one process.
    """
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    saved = {n: m for n, m in sys.modules.items() if n == "mmedited" or n.startswith("mmedited.")}
    for name in saved:
        del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    importlib.invalidate_caches()
    try:
        net = importlib.import_module(
            "mmedited.models.backbones.sr_backbones.ciaosr_net")
        csnln = importlib.import_module("mmedited.models.common.arch_csnln")
        mlp = importlib.import_module(
            "mmedited.models.components.refiners.mlp_refiner")
        restorer = importlib.import_module("mmedited.models.restorers.ciaosr")
        swinir = importlib.import_module("mmedited.models.backbones.sr_backbones.swinir_net")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for name in [n for n in sys.modules if n == "mmedited" or n.startswith("mmedited.")]:
            del sys.modules[name]
        sys.modules.update(saved)
        importlib.invalidate_caches()
    assert net.__file__.startswith(REFERENCE_ROOT), net.__file__
    _REF = types.SimpleNamespace(net=net, csnln=csnln, mlp=mlp, swinir=swinir,
                                 restorer=restorer, make_coord=make_coord)
    return _REF


def build_reference_generator(kind="edsr", mid_channels=64, hidden=(256, 256, 256, 256),
                              num_blocks=2, eval_bsize=None, local_size=2,
                              non_local_attn=True, softmax_scale=1, num_layers=2):
    """Instantiate the reference's generator class with a small encoder."""
    ref = import_reference()
    mlp_cfg = lambda: dict(type="MLPRefiner", in_dim=4, out_dim=3,
                           hidden_list=list(hidden))
    if kind == "edsr":
        cls = ref.net.LocalImplicitSREDSR
        enc = dict(type="EDSR", in_channels=3, out_channels=3,
                   mid_channels=mid_channels, num_blocks=num_blocks)
    elif kind == "rdn":
        cls = ref.net.LocalImplicitSRRDN
        enc = dict(type="RDN", in_channels=3, out_channels=3,
                   mid_channels=mid_channels, num_blocks=num_blocks,
                   upscale_factor=4, num_layers=num_layers, channel_growth=mid_channels)
    else:
        raise ValueError(kind)
    gen = cls(encoder=enc, imnet_q=mlp_cfg(), imnet_k=mlp_cfg(), imnet_v=mlp_cfg(),
              local_size=local_size, feat_unfold=True, eval_bsize=eval_bsize,
              non_local_attn=non_local_attn, softmax_scale=softmax_scale)
    return gen.eval()


def build_reference_swinir_generator(embed_dim=24, depths=(2, 2), num_heads=(2, 2), window_size=4, img_size=8,
                                     mlp_ratio=2, hidden=(16, 16), non_local_attn=False):
    """The reference's LocalImplicitSRSWINIR over its own SwinIR, small, on the CPU."""
    ref = import_reference()
    mlp_cfg = lambda: dict(type="MLPRefiner", in_dim=4, out_dim=3, hidden_list=list(hidden))
    enc = dict(type=ref.swinir.SwinIR, upscale=4, in_chans=3, img_size=img_size, window_size=window_size,
               img_range=1., depths=list(depths), embed_dim=embed_dim, num_heads=list(num_heads),
               mlp_ratio=mlp_ratio, upsampler="pixelshuffle", resi_connection="1conv")
    saved = nn.Module.cuda
    nn.Module.cuda = lambda self, *a, **k: self
    try:
        gen = ref.net.LocalImplicitSRSWINIR(window_size=window_size, encoder=enc, imnet_q=mlp_cfg(),
                                            imnet_k=mlp_cfg(), imnet_v=mlp_cfg(), feat_unfold=True,
                                            eval_bsize=None, non_local_attn=non_local_attn)
    finally:
        nn.Module.cuda = saved
    return gen.eval()
