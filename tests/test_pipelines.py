"""Coordinate / cell generation (ciaosr_b200/pipelines.py) against outputs of the reference's own classes
(oracle/make_pipeline_golden.py -> tests/golden/coord_pipeline.npz), with numpy's RNG seeded the same way."""
import json
import os

import numpy as np
import torch

from ciaosr_b200 import pipelines
from ciaosr_b200.builder import build
from oracle.make_pipeline_golden import inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coord_pipeline.npz")


def test_matches_reference_classes():
    z = np.load(GOLD)
    cases = json.loads(bytes(z["meta"]).decode())
    assert len(cases) == 6
    for i, case in enumerate(cases):
        np.random.seed(case["seed"])
        out = getattr(pipelines, case["cls"])(**case["kw"])(inputs(case))
        want = {k[len(f"{i}_"):]: z[k] for k in z.files if k.startswith(f"{i}_")}
        got = {k: v for k, v in out.items() if torch.is_tensor(v)}
        assert sorted(got) == sorted(want), (case, sorted(got), sorted(want))
        for k, v in want.items():
            assert got[k].shape == v.shape and np.array_equal(got[k].numpy(), v), (case, k)


def test_registered_name_and_test_mode():
    step = build(dict(type="GenerateCoordinateAndCell", scale=4))          # as in configs/001_*rdn*.py:93,115
    out = step(dict(lq=torch.zeros(3, 5, 7)))
    assert out["coord"].shape == (20 * 28, 2) and out["cell"].shape == (20 * 28, 2)
    assert torch.allclose(out["cell"][0], torch.tensor([2 / 20, 2 / 28]))
    np.random.seed(0)
    out = build(dict(type="GenerateCoordinateAndCell", sample_quantity=9))(dict(gt=torch.rand(3, 6, 6)))
    assert out["gt"].shape == (9, 3) and out["coord"].shape == (9, 2)
    assert "GenerateCoordinateAndCell" in repr(step)
