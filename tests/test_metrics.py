"""PSNR / SSIM (host side of the restorer's `evaluate`, basic_restorer.py:101-124) against values computed by
the reference's own metric functions (oracle/make_metrics_golden.py -> tests/golden/metrics.npz)."""
import json
import os

import numpy as np
import pytest
import torch

from ciaosr_b200 import metrics

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")


def _cases():
    z = np.load(GOLD)
    return z["gt"], z["out"], json.loads(bytes(z["meta"]).decode())


def test_psnr_ssim_match_reference_values():
    gt, out, cases = _cases()
    assert len(cases) == 4
    for c in cases:
        p = metrics.psnr(out, gt, c["crop_border"], convert_to=c["convert_to"])
        s = metrics.ssim(out, gt, c["crop_border"], convert_to=c["convert_to"])
        assert abs(p - c["psnr"]) < 1e-6, (c, p)
        assert abs(s - c["ssim"]) < 1e-7, (c, s)


def test_tensor2img_and_identity():
    gt, _, _ = _cases()
    t = torch.from_numpy(gt[..., ::-1].copy()).permute(2, 0, 1).float().div(255.0)[None]     # BGR uint8 -> RGB float
    img = metrics.tensor2img(t)
    assert img.dtype == np.uint8 and np.array_equal(img, gt)
    assert metrics.psnr(img, gt) == float("inf")
    assert abs(metrics.ssim(img, gt) - 1.0) < 1e-12
    with pytest.raises(ValueError):
        metrics.psnr(img, gt, convert_to="hsv")


def test_device_metrics_match_host_metrics():
    """The torch restatement used by `evaluate` for CUDA tensors (runs on any device; here CPU) reproduces the
    host metrics -- and therefore the reference's values -- on the golden pair."""
    gt, out, cases = _cases()
    to_t = lambda img: torch.from_numpy(img[..., ::-1].copy()).permute(2, 0, 1).float().div(255.0)[None]
    tg, to = to_t(gt), to_t(out)
    for c in cases:
        p = metrics.psnr_device(to, tg, c["crop_border"], convert_to=c["convert_to"])
        s = metrics.ssim_device(to, tg, c["crop_border"], convert_to=c["convert_to"])
        assert abs(p - c["psnr"]) < 1e-5, (c, p)
        assert abs(s - c["ssim"]) < 1e-7, (c, s)
