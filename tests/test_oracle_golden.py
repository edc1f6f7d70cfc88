"""The CPU oracle against the golden vectors minted from the reference
(oracle/make_golden.py).  This is what pins the oracle (the
reference itself has no tests or fixtures for the path)."""
import pytest
import torch

from oracle import ciaosr_oracle as orc
from tests.util import CSATTN_CASES, HEAD_CASES, build_generator, head_weights, load_case, max_abs

# fp32 reassociation only: matmul vs conv formulations of the same sums
TOL = 2e-5


@pytest.mark.parametrize("name", CSATTN_CASES)
def test_cross_scale_attention(name):
    meta, a = load_case(name)
    holder = torch.nn.Module()
    from ciaosr_b200.cross_scale_attention import CrossScaleAttention
    from ciaosr_b200 import synth
    holder.cs_attn = CrossScaleAttention(channel=meta["c"], scale=[2])
    synth.fill_module(holder, meta["seed"])
    w = {k: v.detach() for k, v in holder.state_dict().items()}
    out = orc.cross_scale_attention(a["feature"], w)
    assert out.shape == a["out"].shape
    assert max_abs(out, a["out"]) < TOL


@pytest.mark.parametrize("name", HEAD_CASES)
def test_head(name):
    meta, a = load_case(name)
    w = head_weights(build_generator(meta))
    kw = dict(local_size=meta["local_size"], non_local_attn=meta["non_local"])
    if meta["non_local"]:
        nl = orc.cross_scale_attention(a["feature"], w)
        assert max_abs(nl, a["nonlocal"]) < TOL
    for tag in meta["tags"]:
        coord, cell = a[f"coord_{tag}"], a[f"cell_{tag}"]
        pred = orc.query_rgb(a["feature"], coord, cell, w, **kw)
        assert max_abs(pred, a[f"pred_{tag}"]) < TOL, tag
        out = orc.head_forward(a["x_lr"], a["feature"], coord, cell, w,
                               eval_bsize=meta["eval_bsize"], **kw)
        assert max_abs(out, a[f"out_{tag}"]) < TOL, tag


def test_make_coord_matches_fixture():
    meta, a = load_case("head_small")
    th, tw = meta["h"] * 2, meta["w"] * 2
    assert torch.equal(orc.make_coord((th, tw)), a["coord_s2"][0])
    assert torch.equal(orc.cell_for((th, tw), th * tw), a["cell_s2"][0])


@pytest.mark.parametrize("name", ["clip_small", "clip_real_small"])      # CiaoSR.clip_test / RealCiaoSR.clip_test
def test_clip_test(name):
    meta, a = load_case(name)
    g = build_generator(dict(c=meta["c"], hidden=meta["hidden"], eval_bsize=meta["eval_bsize"],
                             seed=meta["seed"], non_local=meta.get("non_local", True)))
    w = head_weights(g)

    def model(patch, coord, cell):
        with torch.no_grad():
            feat = g.gen_feature(patch)[0]
        return orc.head_forward(patch, feat, coord, cell, w, eval_bsize=meta["eval_bsize"],
                                non_local_attn=meta.get("non_local", True))

    out = orc.clip_test(a["lq"], model, meta["scale"], meta["tile"], meta["overlap"])
    assert out.shape == a["out"].shape
    assert max_abs(out, a["out"]) < TOL


def test_flop_model():
    # 8 865 280 FLOP per output pixel at C=64 with the non-local branch
    assert orc.head_flops_per_query(64) == 8865280
    assert orc.head_flops_per_query(180) == 18486784
    assert orc.head_flops_per_query(180, non_local=False) == 17657344


# ---- BASELINE-size goldens (oracle/make_golden_full.py) --------------------------------------------------------
def test_head_full_frac():
    """Fractional scales at the real head dimensions (C = 64, hidden 256x4)."""
    meta, a = load_case("full_frac")
    from ciaosr_b200 import synth
    w = head_weights(build_generator(meta))
    feat = synth.synth_feature(meta["b"], 64, meta["h"], meta["w"], meta["seed"])
    x_lr = synth.synth_lr_image(meta["b"], meta["h"], meta["w"], meta["seed"])
    for tag in meta["tags"]:
        out = orc.head_forward(x_lr, feat, a[f"coord_{tag}"], a[f"cell_{tag}"], w, eval_bsize=meta["eval_bsize"])
        assert max_abs(out, a[f"out_{tag}"]) < TOL, tag


def test_head_config2_crop():
    """BASELINE.json config 2, one crop: the oracle head on the reference encoder's own feature map (stored with
    the golden) against the reference's end-to-end output: pins the oracle at the size bench.py times."""
    import bench
    meta, a = load_case("full_cfg2")
    m = bench.build_model()
    w = {k: v.detach() for k, v in m.generator.state_dict().items()}
    lq, coord, cell = bench.make_inputs(16, meta["seed"])
    lq = lq - torch.tensor(bench.RGB_MEAN).view(1, 3, 1, 1)
    out = orc.head_forward(lq[:1], a["feature0"][None], coord[:1], cell[:1], w, eval_bsize=bench.EVAL_BSIZE)
    assert max_abs(out, a["out"][:1]) < TOL
