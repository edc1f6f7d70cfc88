"""PSNR / SSIM on uint8-rounded images, optionally on the Y channel, as the reference
evaluates them (mmedited/core/evaluation/metrics.py:181-318 via basic_restorer.py:101-124).
Host-side numpy, like the reference; not on the accelerated path."""
import numpy as np
import torch


def tensor2img(tensor, min_max=(0, 1)):
    """[1,3,H,W] or [3,H,W] RGB float tensor -> HWC BGR uint8 (mmedit.core.tensor2img)."""
    t = tensor.squeeze(0).float().detach().cpu().clamp_(*min_max)
    t = (t - min_max[0]) / (min_max[1] - min_max[0])
    img = t.numpy()
    if img.ndim == 3:
        img = np.transpose(img[[2, 1, 0], :, :], (1, 2, 0))
    return (img * 255.0).round().astype(np.uint8)


def _to_y(img):
    """BGR [0,255] -> Y [0,255] (mmcv.bgr2ycbcr(y_only=True) on img/255, times 255)."""
    img = img.astype(np.float32) / 255.0
    y = np.dot(img, [24.966, 128.553, 65.481]) + 16.0
    return (y / 255.0).astype(np.float32) * 255.0          # mmcv returns float32 for a float32 input


def _prep(img1, img2, crop_border, convert_to):
    assert img1.shape == img2.shape, f"Image shapes are different: {img1.shape}, {img2.shape}."
    img1, img2 = img1.astype(np.float32), img2.astype(np.float32)
    if isinstance(convert_to, str) and convert_to.lower() == "y":
        img1, img2 = _to_y(img1)[..., None], _to_y(img2)[..., None]
    elif convert_to is not None:
        raise ValueError('Wrong color model. Supported values are "Y" and None.')
    if crop_border != 0:
        img1 = img1[crop_border:-crop_border, crop_border:-crop_border, ...]
        img2 = img2[crop_border:-crop_border, crop_border:-crop_border, ...]
    return img1, img2


def psnr(img1, img2, crop_border=0, input_order="HWC", convert_to=None):
    img1, img2 = _prep(img1, img2, crop_border, convert_to)
    mse = np.mean((img1 - img2) ** 2)
    return float("inf") if mse == 0 else float(20.0 * np.log10(255.0 / np.sqrt(mse)))


def _gauss_kernel(size=11, sigma=1.5):
    ax = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    k = np.exp(-(ax ** 2) / (2 * sigma ** 2))
    k /= k.sum()
    return np.outer(k, k)


def _filter_valid(img, window):
    t = torch.from_numpy(img)[None, None].double()
    w = torch.from_numpy(window)[None, None]
    return torch.nn.functional.conv2d(t, w)[0, 0].numpy()


def _ssim(img1, img2):
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    img1, img2 = img1.astype(np.float64), img2.astype(np.float64)
    win = _gauss_kernel()
    mu1, mu2 = _filter_valid(img1, win), _filter_valid(img2, win)
    s1 = _filter_valid(img1 ** 2, win) - mu1 ** 2
    s2 = _filter_valid(img2 ** 2, win) - mu2 ** 2
    s12 = _filter_valid(img1 * img2, win) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2))
    return m.mean()


def ssim(img1, img2, crop_border=0, input_order="HWC", convert_to=None):
    """metrics.py:264-318.  Reference quirk kept on purpose: its crop is ``img[c:-c, c:-c, None]``, which for a
    3-channel image yields [h, w, 1, 3], so the per-channel loop runs once and scores channel 0 (B of BGR) only;
    with ``convert_to='y'`` (what every reference config uses together with ``crop_border``) it is the Y channel
    either way."""
    img1, img2 = _prep(img1, img2, crop_border, convert_to)
    if crop_border != 0 and img1.shape[2] > 1:
        img1, img2 = img1[..., :1], img2[..., :1]
    return float(np.mean([_ssim(img1[..., i], img2[..., i]) for i in range(img1.shape[2])]))


# ---- the same metrics evaluated where the image already is ------------------------------------
# `BasicRestorer.evaluate` in the reference moves both frames to the host (tensor2img) and scores them in numpy /
# cv2; for a 4K / 8K frame that transfer dominates.  The functions below reproduce the same arithmetic with torch
# ops on the tensor's own device in float64 (uint8 rounding, BGR order, mmcv's Y conversion incl. its float32
# rounding step, the crop and its channel-0 quirk, the 11x11 sigma-1.5 valid-window SSIM) and return Python floats.
def _quantise_bgr(t):
    """[1,3,H,W] / [3,H,W] RGB float in [0,1] -> [H,W,3] BGR, uint8-rounded, float64."""
    t = t.squeeze(0).detach().float().clamp(0, 1)
    return (t.flip(0).permute(1, 2, 0) * 255.0).round().double()


def _prep_device(out, gt, crop_border, convert_to):
    a, b = _quantise_bgr(out), _quantise_bgr(gt)
    assert a.shape == b.shape, f"Image shapes are different: {tuple(a.shape)}, {tuple(b.shape)}."
    if isinstance(convert_to, str) and convert_to.lower() == "y":
        coef = torch.tensor([24.966, 128.553, 65.481], dtype=torch.float64, device=a.device)

        def to_y(x):
            x32 = (x.float() / 255.0).double()                       # img.astype(float32) / 255.
            return (((x32 @ coef) + 16.0) / 255.0).float().double().unsqueeze(-1) * 255.0
        a, b = to_y(a), to_y(b)
    elif convert_to is not None:
        raise ValueError('Wrong color model. Supported values are "Y" and None.')
    if crop_border != 0:
        a = a[crop_border:-crop_border, crop_border:-crop_border]
        b = b[crop_border:-crop_border, crop_border:-crop_border]
    return a, b


def psnr_device(out, gt, crop_border=0, convert_to=None):
    a, b = _prep_device(out, gt, crop_border, convert_to)
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse == 0 else float(20.0 * np.log10(255.0 / np.sqrt(mse)))


def ssim_device(out, gt, crop_border=0, convert_to=None):
    a, b = _prep_device(out, gt, crop_border, convert_to)
    if crop_border != 0 and a.shape[2] > 1:                          # the reference's crop quirk, see ssim()
        a, b = a[..., :1], b[..., :1]
    win = torch.from_numpy(_gauss_kernel()).to(a.device)[None, None]
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    x, y = a.permute(2, 0, 1).unsqueeze(1), b.permute(2, 0, 1).unsqueeze(1)      # [C,1,H,W]
    f = lambda t: torch.nn.functional.conv2d(t, win)
    mu1, mu2 = f(x), f(y)
    s1, s2, s12 = f(x * x) - mu1 ** 2, f(y * y) - mu2 ** 2, f(x * y) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2))
    return float(m.flatten(1).mean(1).mean())
