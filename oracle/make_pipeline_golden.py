"""Golden outputs of the reference's coordinate / cell pipeline step, produced by RUNNING its classes
(mmedited/datasets/pipelines/generate_assistant.py, imported in place through oracle/ref_harness.py's stubs)
with numpy's RNG seeded.   python -m oracle.make_pipeline_golden   -> tests/golden/coord_pipeline.npz"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

CASES = [
    dict(cls="GenerateCoordinateAndCell1", kw=dict(sample_quantity=37), keys=["gt", "gt_unsharp"], shape=[3, 9, 11], seed=1),
    dict(cls="GenerateCoordinateAndCell1", kw=dict(sample_quantity=20, is_shuffle=False), keys=["gt", "gt_unsharp"], shape=[3, 8, 8], seed=2),
    dict(cls="GenerateCoordinateAndCell1", kw=dict(sample_quantity=20), keys=["gt"], shape=[3, 6, 7], seed=3),     # no unsharp: no sampling
    dict(cls="GenerateCoordinateAndCell1", kw=dict(scale=2.5), keys=["lq"], shape=[3, 5, 6], seed=4),
    dict(cls="GenerateCoordinateAndCell1", kw=dict(target_size=[7, 9]), keys=[], shape=[3, 1, 1], seed=5),
    dict(cls="GenerateCoordinateAndCell2", kw=dict(sample_quantity=15, scale=4, scale1=3), keys=["gt"], shape=[3, 12, 16], seed=6),
]


def reference_module():
    rh.import_reference()                               # installs the mmcv / mmedit stubs
    saved = {n: m for n, m in sys.modules.items() if n == "mmedited" or n.startswith("mmedited.")}
    for n in saved:
        del sys.modules[n]
    sys.path.insert(0, rh.REFERENCE_ROOT)
    importlib.invalidate_caches()
    try:
        mod = importlib.import_module("mmedited.datasets.pipelines.generate_assistant")
    finally:
        sys.path.remove(rh.REFERENCE_ROOT)
        for n in [n for n in sys.modules if n == "mmedited" or n.startswith("mmedited.")]:
            del sys.modules[n]
        sys.modules.update(saved)
        importlib.invalidate_caches()
    assert mod.__file__.startswith(rh.REFERENCE_ROOT)
    return mod


def inputs(case):
    rs = np.random.RandomState(100 + case["seed"])
    return {k: torch.from_numpy(rs.uniform(0, 1, size=case["shape"]).astype(np.float32)) for k in case["keys"]}


def main():
    mod = reference_module()
    arrays = {}
    for i, case in enumerate(CASES):
        np.random.seed(case["seed"])
        out = getattr(mod, case["cls"])(**case["kw"])(inputs(case))
        for k, v in out.items():
            if torch.is_tensor(v):
                arrays[f"{i}_{k}"] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "coord_pipeline.npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(CASES).encode(), dtype=np.uint8), **arrays)
    print(path, sorted(arrays))


if __name__ == "__main__":
    main()
