"""Training support of the head: the differentiable recompute used by the backward
(ciaosr_b200/head_autograd.py) against the value and the gradients the UNMODIFIED reference's autograd graph
produced (tests/golden/train_small.npz, oracle/make_golden.py::train_case), and -- on a GPU -- the whole
``generator(lq, coord, cell)`` training forward / backward and ``CiaoSR.train_step`` through the native forward."""
import pytest
import torch

from ciaosr_b200 import head_autograd, synth
from tests.util import build_generator, load_case, max_abs


def _setup(dev="cpu"):
    meta, a = load_case("train_small")
    g = build_generator(meta, dev).train()
    feat = synth.synth_feature(meta["b"], meta["c"], meta["h"], meta["w"], meta["seed"]).to(dev)
    x_lr = synth.synth_lr_image(meta["b"], meta["h"], meta["w"], meta["seed"]).to(dev)
    return meta, {k: v.to(dev) for k, v in a.items()}, g, feat, x_lr


def _rel(a, b):
    return max_abs(a, b) / max(float(b.abs().max()), 1e-12)


def test_recompute_matches_reference_value_and_gradients():
    """CPU: the torch restatement alone (what HeadFunction.backward differentiates)."""
    meta, a, g, feat, x_lr = _setup()
    params = g._head_params_live()
    feat.requires_grad_(True)
    pred = head_autograd.query_rgb(feat, a["coord"], a["cell"], params, local_size=meta["local_size"])
    res = torch.nn.functional.grid_sample(x_lr, a["coord"].flip(-1).unsqueeze(1), mode="bilinear",
                                          padding_mode="border", align_corners=False)[:, :, 0, :].permute(0, 2, 1)
    assert max_abs(pred + res, a["pred"]) < 2e-5
    loss = (pred + res - a["gt"]).abs().mean()
    assert abs(float(loss) - float(a["loss"])) < 1e-6
    loss.backward()
    assert _rel(feat.grad, a["grad_feature"]) < 1e-4
    checked = 0
    for k, p in params.items():
        if "grad__" + k in a:
            assert _rel(p.grad, a["grad__" + k]) < 1e-4, k
            checked += 1
    assert checked >= 20          # 3 MLPs x 3 layers x (w, b) + the cross-scale attention's convolutions


@pytest.mark.gpu
def test_training_forward_backward_on_gpu():
    """GPU: value from the native kernels, gradient from the recompute, through the generator's own forward."""
    dev = torch.device("cuda:0")
    meta, a, g, feat, x_lr = _setup(dev)
    feat.requires_grad_(True)
    g.gen_feature = lambda _x: [feat]
    pred = g(x_lr, a["coord"], a["cell"])                       # test_mode=False, grad enabled
    assert pred.requires_grad and max_abs(pred.detach().cpu(), a["pred"].cpu()) < 1e-4
    loss = (pred - a["gt"]).abs().mean()
    loss.backward()
    assert _rel(feat.grad.cpu(), a["grad_feature"].cpu()) < 1e-3
    for k, p in g._head_params_live().items():
        if "grad__" + k in a:
            assert _rel(p.grad.cpu(), a["grad__" + k].cpu()) < 1e-3, k


@pytest.mark.gpu
def test_train_step_reduces_the_loss():
    """CiaoSR.train_step (ciaosr.py:60-109) end to end: encoder under autograd + native head; a few Adam steps on
    one batch reduce the pixel loss, and the inference forward sees the updated weights (plans are rebuilt)."""
    from ciaosr_b200.builder import build
    from ciaosr_b200.restorers import CiaoSR
    from tests.util import generator_cfg
    dev = torch.device("cuda:0")
    m = build(dict(type=CiaoSR, generator=generator_cfg(16, [32, 32], None), pixel_loss=dict(type="L1Loss"),
                   rgb_mean=(0.4488, 0.4371, 0.4040), rgb_std=(1., 1., 1.)))
    synth.fill_module(m.generator, 4)
    m = m.train().to(dev)
    gen = torch.Generator().manual_seed(0)
    b, h, w, nq = 2, 12, 12, 128
    batch = dict(lq=torch.rand(b, 3, h, w, generator=gen).to(dev), gt=torch.rand(b, nq, 3, generator=gen).to(dev),
                 coord=(torch.rand(b, nq, 2, generator=gen) * 2 - 1).to(dev),
                 cell=torch.full((b, nq, 2), 2 / (h * 2.0)).to(dev))
    opt = torch.optim.Adam(m.generator.parameters(), lr=1e-4)
    losses = [m.train_step(batch, opt)["log_vars"]["loss_pix"] for _ in range(6)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    with torch.no_grad():
        out = m.generator(batch["lq"] - 0.44, batch["coord"], batch["cell"], test_mode=True)
    assert torch.isfinite(out).all()
