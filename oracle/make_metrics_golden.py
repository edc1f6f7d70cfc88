"""Golden values for the evaluation metrics, computed by the REFERENCE's own `psnr` / `ssim`
(mmedited/core/evaluation/metrics.py:181-318), extracted from the file by name and executed with the real
cv2 / numpy and a 10-line restatement of the one mmcv function they call (`mmcv.bgr2ycbcr`, mmcv 1.x
image/colorspace.py: float32 input in [0,1] -> Y = (24.966 B + 128.553 G + 65.481 R + 16) / 255).

    python -m oracle.make_metrics_golden        # rewrites tests/golden/metrics.npz   (build container only)
"""
import ast
import json
import os
import sys
import types

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/mmedited/core/evaluation/metrics.py"


def _bgr2ycbcr(img, y_only=False):
    assert y_only and img.dtype == np.float32
    return ((np.dot(img, [24.966, 128.553, 65.481]) + 16.0) / 255.0).astype(np.float32)


def reference_metrics():
    tree = ast.parse(open(SRC).read())
    want = {"psnr", "ssim", "_ssim", "reorder_image"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"np": np, "cv2": cv2, "mmcv": types.SimpleNamespace(bgr2ycbcr=_bgr2ycbcr)}
    exec(compile(ast.Module(body=body, type_ignores=[]), SRC, "exec"), ns)
    return ns["psnr"], ns["ssim"]


def main():
    psnr, ssim = reference_metrics()
    rs = np.random.RandomState(5)
    h, w = 48, 40
    gt = (rs.uniform(0, 1, size=(h, w, 3)) * 255).round().astype(np.uint8)
    out = np.clip(gt.astype(np.float32) + rs.normal(0, 6.0, size=gt.shape), 0, 255).round().astype(np.uint8)
    cases = []
    for crop in (0, 4):
        for conv in (None, "y"):
            cases.append(dict(crop_border=crop, convert_to=conv, psnr=float(psnr(out, gt, crop, convert_to=conv)),
                              ssim=float(ssim(out, gt, crop, convert_to=conv))))
    path = os.path.join(ROOT, "tests", "golden", "metrics.npz")
    np.savez_compressed(path, gt=gt, out=out, meta=np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8))
    print(path, cases)


if __name__ == "__main__":
    sys.exit(main())
