"""Host-side logic of the boundary (CPU only): config loading, builders, state_dict
layout, constructor side effects and error behaviour of the reference API."""
import os

import pytest
import torch

from ciaosr_b200 import make_coord
from ciaosr_b200.builder import Config, build, build_loss
from ciaosr_b200.generators import LocalImplicitSREDSR, LocalImplicitSRRDN
from ciaosr_b200.restorers import CiaoSR
from oracle import ciaosr_oracle as orc
from tests.util import generator_cfg

REF_CONFIGS = "/root/reference/configs"


def test_make_coord_same_as_oracle():
    for shape in [(7, 5), (192, 192), (3,)]:
        assert torch.equal(make_coord(shape), orc.make_coord(shape))
    assert make_coord((4, 6), flatten=False).shape == (4, 6, 2)


def test_ctor_mutates_config_like_reference():
    # ciaosr_net.py:61-76 rewrites the imnet dicts in place
    cfg = generator_cfg(64, (256, 256, 256, 256))
    g = build(cfg)
    assert cfg["imnet_k"]["in_dim"] == 580 and cfg["imnet_k"]["out_dim"] == 576
    assert cfg["imnet_v"]["in_dim"] == 644 and cfg["imnet_v"]["out_dim"] == 640
    assert cfg["imnet_q"]["in_dim"] == 640 and cfg["imnet_q"]["out_dim"] == 3
    assert not hasattr(g, "encoder")      # hoisted and deleted, ciaosr_net.py:388-391
    n = sum(p.numel() for n_, p in g.named_parameters() if n_.startswith("imnet_"))
    assert n == 494144 + 526976 + 362243


def test_state_dict_keys():
    g = build(generator_cfg(16, (32, 32)))
    keys = set(g.state_dict().keys())
    for k in ["imnet_k.layers.0.weight", "imnet_k.layers.4.bias", "cs_attn.conv_match_1.0.weight",
              "cs_attn.conv_match_1.1.weight", "cs_attn.conv_assembly.0.bias", "cs_attn.down.weight",
              "cs_attn.escape_NaN", "conv_first.weight", "body.0.conv1.weight",
              "conv_after_body.bias"]:
        assert k in keys, k
    rdn = build(dict(type=LocalImplicitSRRDN,
                     encoder=dict(type="RDN", in_channels=3, out_channels=3, mid_channels=16,
                                  num_blocks=2, upscale_factor=4, num_layers=2, channel_growth=16),
                     **{k: dict(type="MLPRefiner", in_dim=4, out_dim=3, hidden_list=[8])
                        for k in ("imnet_q", "imnet_k", "imnet_v")}))
    keys = set(rdn.state_dict().keys())
    for k in ["sfe1.weight", "sfe2.bias", "rdbs.1.layers.0.conv.weight", "rdbs.0.lff.weight",
              "gff.0.weight", "gff.1.bias"]:
        assert k in keys, k


def test_init_weights_errors():
    g = build(generator_cfg(8, (8,)))
    g.init_weights(None)
    with pytest.raises(TypeError, match='"pretrained" must be a str or None'):
        g.init_weights(123)


def test_no_cpu_path():
    g = build(generator_cfg(8, (8,))).eval()
    x = torch.zeros(1, 3, 6, 6)
    coord = make_coord((12, 12)).unsqueeze(0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g(x, coord, torch.ones_like(coord), test_mode=True)


def test_training_forward_has_no_cpu_path_either():
    """The training forward takes its value from the native kernels too (head_autograd.HeadFunction); on CPU
    tensors it refuses like the inference forward."""
    g = build(generator_cfg(8, (8,))).train()
    coord = make_coord((4, 4)).unsqueeze(0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g(torch.zeros(1, 3, 4, 4), coord, torch.ones_like(coord))


def test_tile_origins_match_oracle():
    for n, t, o in [(40, 24, 8), (256, 192, 32), (192, 192, 32), (720, 192, 32), (1080, 128, 32)]:
        assert CiaoSR.tile_origins(n, t, o) == orc.tile_origins(n, t, o)
    assert len(CiaoSR.tile_origins(1280, 192, 32)) * len(CiaoSR.tile_origins(720, 192, 32)) == 40


def test_l1_loss():
    l = build_loss(dict(type="L1Loss", loss_weight=2.0, reduction="mean"))
    assert float(l(torch.ones(4), torch.zeros(4))) == 2.0


def test_local_config_file(tmp_path):
    p = tmp_path / "cfg.py"
    p.write_text(
        "from mmedited.models.restorers.ciaosr import CiaoSR\n"
        "from mmedited.models.backbones.sr_backbones.ciaosr_net import LocalImplicitSREDSR\n"
        "val_scale = 2\n"
        "model = dict(type=CiaoSR, generator=dict(type=LocalImplicitSREDSR,\n"
        "    encoder=dict(type='EDSR', in_channels=3, out_channels=3, mid_channels=8, num_blocks=1),\n"
        "    imnet_q=dict(type='MLPRefiner', in_dim=4, out_dim=3, hidden_list=[8]),\n"
        "    imnet_k=dict(type='MLPRefiner', in_dim=4, out_dim=3, hidden_list=[8]),\n"
        "    imnet_v=dict(type='MLPRefiner', in_dim=4, out_dim=3, hidden_list=[8]),\n"
        "    feat_unfold=True, eval_bsize=30000),\n"
        "  rgb_mean=(0.4488, 0.4371, 0.4040), rgb_std=(1., 1., 1.),\n"
        "  pixel_loss=dict(type='L1Loss', loss_weight=1.0, reduction='mean'))\n"
        "test_cfg = dict(metrics=['PSNR'], crop_border=val_scale, scale=val_scale, tile=8, tile_overlap=2)\n")
    cfg = Config.fromfile(str(p))
    m = build(cfg.model, test_cfg=cfg.test_cfg)
    assert isinstance(m, CiaoSR) and m.test_cfg["tile"] == 8
    assert m.generator.eval_bsize == 30000


@pytest.mark.skipif(not os.path.isdir(REF_CONFIGS), reason="reference configs not mounted")
@pytest.mark.parametrize("name", [
    "001_localimplicitsr_rdn_div2k_g1_c64b16_1000k_unfold_lec_mulwkv_res_nonlocal.py"])
def test_reference_configs_load_unchanged(name):
    """The RDN config (BASELINE.json config 2) is the only 001 config that parses: the
    EDSR and SwinIR 001 files ship with an unclosed `data = dict(` (SyntaxError in the
    reference itself, see test_reference_edsr_config_is_broken_upstream)."""
    cfg = Config.fromfile(os.path.join(REF_CONFIGS, name))
    model = build(cfg.model, test_cfg=cfg.test_cfg)
    assert isinstance(model, CiaoSR)
    assert model.generator.eval_bsize == 30000 and model.generator.imnet_dim == 64
    assert model.test_cfg["tile"] == 192 and model.test_cfg["tile_overlap"] == 32
    assert tuple(model.lq_mean.flatten().tolist()) == pytest.approx((0.4488, 0.4371, 0.4040))


@pytest.mark.skipif(not os.path.isdir(REF_CONFIGS), reason="reference configs not mounted")
def test_reference_edsr_config_is_broken_upstream():
    name = "001_localimplicitsr_edsr_div2k_g1_c64b16_1000k_unfold_lec_mulwkv_res_nonlocal.py"
    with pytest.raises(SyntaxError):
        compile(open(os.path.join(REF_CONFIGS, name)).read(), name, "exec")


def test_native_state_lives_outside_the_module():
    """plans / graphs hold ctypes structs with pointers; they are cached outside the nn.Module so that
    copy.deepcopy and pickle of a generator that has already run keep working, and a copy starts with no cache."""
    import copy
    import ctypes
    import pickle
    from ciaosr_b200 import _lib, native
    from tests.util import build_generator
    g = build_generator(dict(c=8, hidden=[16], eval_bsize=None, local_size=2, non_local=True, seed=1))
    desc = _lib.HeadDesc()                                       # what a live HeadPlan carries
    desc.imnet_q.weight[0] = ctypes.addressof(ctypes.c_float(1.0))
    native.module_cache(g)["plan"] = desc
    native.module_cache(g.imnet_q.layers[0])["linear_plan"] = ("key", desc)
    assert g._plan is desc
    g2 = copy.deepcopy(g)
    assert g2._plan is None and not native.module_cache(g2.imnet_q.layers[0])
    g3 = pickle.loads(pickle.dumps(g))
    assert g3._plan is None
    assert all(torch.equal(a, b) for a, b in zip(g.state_dict().values(), g2.state_dict().values()))
