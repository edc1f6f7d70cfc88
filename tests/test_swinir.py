"""SwinIR trunk (host PyTorch encoder of configs 001-swinir / 002) against the golden features the
reference's own SwinIR produced (oracle/make_golden.py::swinir_case), checkpoint-key compatibility,
and the 002 real-world configs loading unchanged.  CPU only: the encoder is host code."""
import os

import pytest
import torch

from ciaosr_b200 import synth
from ciaosr_b200.builder import Config, build
from ciaosr_b200.generators import LocalImplicitSRSWINIR
from ciaosr_b200.restorers import RealCiaoSR
from ciaosr_b200.swinir import SwinIR, shifted_window_mask
from tests.util import load_case, max_abs

REF_CONFIGS = "/root/reference/configs"


def _generator(cfg, hidden=(16, 16), non_local=False):
    mlp = lambda: dict(type="MLPRefiner", in_dim=4, out_dim=3, hidden_list=list(hidden))
    enc = dict(type=SwinIR, upscale=4, in_chans=3, img_size=cfg["img_size"], window_size=cfg["window_size"],
               img_range=1., depths=list(cfg["depths"]), embed_dim=cfg["embed_dim"],
               num_heads=list(cfg["num_heads"]), mlp_ratio=cfg["mlp_ratio"], upsampler="pixelshuffle",
               resi_connection="1conv")
    return build(dict(type=LocalImplicitSRSWINIR, window_size=cfg["window_size"], encoder=enc, imnet_q=mlp(),
                      imnet_k=mlp(), imnet_v=mlp(), feat_unfold=True, eval_bsize=None,
                      non_local_attn=non_local)).eval()


def test_trunk_matches_reference_golden():
    meta, a = load_case("swinir_trunk")
    g = _generator(meta["cfg"])
    # same state_dict keys and shapes as the reference generator: released checkpoints load with strict=True
    ours = {k: list(v.shape) for k, v in g.state_dict().items()}
    assert ours == meta["state_keys"]
    synth.fill_module(g, meta["seed"])
    with torch.no_grad():
        for tag in meta["tags"]:
            feat = g.gen_feature(a[f"x_{tag}"])[0]
            ref = a[f"feat_{tag}"]
            assert feat.shape == ref.shape
            assert max_abs(feat, ref) < 2e-5 * max(1.0, float(ref.abs().max())), tag


def test_shift_mask_regions():
    m = shifted_window_mask((8, 12), 4, 2)
    assert m.shape == (6, 16, 16) and set(m.unique().tolist()) <= {0.0, -100.0}
    assert float(m[0].abs().sum()) == 0.0            # the top-left window is never cut by the cyclic shift
    assert float(m[-1].abs().sum()) > 0.0            # the bottom-right one mixes four regions


def test_drop_path_and_init_errors():
    enc = SwinIR(img_size=8, window_size=4, embed_dim=12, depths=(2,), num_heads=(2,), mlp_ratio=2, drop_path_rate=0.5,
                 compress_ratio=3, squeeze_factor=30)                   # unknown keywords are accepted like upstream
    assert enc.embed_dim == 12
    with pytest.raises(TypeError):
        enc.init_weights(pretrained=3)
    enc.init_weights(None)
    x = torch.randn(2, 12, 8, 8)
    enc.eval()
    y = enc.layers[0](enc.patch_embed(x), (8, 8))
    assert y.shape == (2, 64, 12) and torch.isfinite(y).all()


@pytest.mark.skipif(not os.path.isdir(REF_CONFIGS), reason="reference configs not mounted")
@pytest.mark.parametrize("name", [
    "002_real_wogan_localimplicitsr_swinir_df2k_g1_c64b16_1000k_unfold_lec_mulwkv.py",
    "002_real_gan_localimplicitsr_swinir_df2k_g1_c64b16_1000k_unfold_lec_mulwkv.py"])
def test_reference_002_configs_load_unchanged(name):
    cfg = Config.fromfile(os.path.join(REF_CONFIGS, name))
    m = build(cfg.model, test_cfg=cfg.test_cfg)
    assert isinstance(m, RealCiaoSR) and m.is_use_ema and m.generator_ema is not None
    g = m.generator
    assert g.imnet_dim == 180 and g.non_local_attn is False and g.res is False and g.window_size == 8
    assert m.test_cfg["tile"] in (128, 256) and m.test_cfg["tile_overlap"] == 32
    keys = m.state_dict().keys()
    assert "generator_ema.layers.5.residual_group.blocks.5.attn.qkv.weight" in keys and "step_counter" in keys
    assert m._test_generator() is m.generator_ema
    with pytest.raises(NotImplementedError):
        cfg.train_pipeline[0] if False else __import__("mmedited.datasets.pipelines.crop", fromlist=["x"]) \
            .PairedRandomCropwScale(gt_patch_size=48)
