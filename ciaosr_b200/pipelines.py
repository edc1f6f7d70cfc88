"""Coordinate / cell generation, the last step of the reference's data pipelines
(mmedited/datasets/pipelines/generate_assistant.py:8-191; string type 'GenerateCoordinateAndCell' in the 001
configs, e.g. configs/001_localimplicitsr_rdn_*.py:75,93,115, resolves to mmedit 0.11's class of that name).

It is the producer of the `coord` / `cell` / flattened `gt` tensors the generator boundary consumes
(SURVEY.md 8f #4).  The classes work on whatever device their input tensors live on -- the coordinate grid is
built there -- so the step can run after the batch has been moved to the GPU; the random query subset uses
numpy's global RNG exactly like the reference (`np.random.choice(n, q, replace=False)`), so a seeded run picks
the same queries.

    GenerateCoordinateAndCell    mmedit 0.11 (not vendored under /root/reference; restated from its published
                                 source -- identical to ...1 below except that sampling only needs 'gt')
    GenerateCoordinateAndCell1   generate_assistant.py:8-102   (real-world configs: samples only when
                                 'gt_unsharp' is present; optional contiguous window instead of a shuffle)
    GenerateCoordinateAndCell2   generate_assistant.py:105-191 (coordinates for scale1 on a GT made for scale)
"""
import numpy as np
import torch

from .coords import make_coord


def _flatten_rgb(img):
    """[3, H, W] -> [H*W, 3] (generate_assistant.py:59-60)."""
    return img.contiguous().view(3, -1).permute(1, 0)


def _cell_for(coord, target_size):
    cell = torch.ones_like(coord)
    cell[:, 0] *= 2 / target_size[-2]
    cell[:, 1] *= 2 / target_size[-1]
    return cell


class GenerateCoordinateAndCell:
    """results['gt'] [3,H,W] (train / val) or results['lq'] + scale, or target_size (test)
    -> results['coord'] [Q,2], results['cell'] [Q,2], results['gt'] [Q,3]."""

    needs_unsharp = False          # GenerateCoordinateAndCell1 samples only when 'gt_unsharp' is present

    def __init__(self, sample_quantity=None, scale=None, target_size=None, is_shuffle=True):
        self.sample_quantity = sample_quantity
        self.scale = scale
        self.target_size = target_size
        self.is_shuffle = is_shuffle

    def _device(self, results):
        for key in ("gt", "lq"):
            if key in results and torch.is_tensor(results[key]):
                return results[key].device
        return torch.device("cpu")

    def __call__(self, results):
        device = self._device(results)
        if "gt" in results:
            self.target_size = results["gt"].shape
            results["gt"] = _flatten_rgb(results["gt"])
            if "gt_unsharp" in results:
                results["gt_unsharp"] = _flatten_rgb(results["gt_unsharp"])
        elif self.scale is not None and "lq" in results:
            _, h_lr, w_lr = results["lq"].shape
            self.target_size = (round(h_lr * self.scale), round(w_lr * self.scale))
        else:
            assert self.target_size is not None
            assert len(self.target_size) >= 2
        coord = make_coord(self.target_size[-2:]).to(device)
        sample = self.sample_quantity is not None and "gt" in results and \
            (not self.needs_unsharp or "gt_unsharp" in results)
        if sample:
            n = len(coord)
            if self.is_shuffle:
                idx = np.random.choice(n, self.sample_quantity, replace=False)
            else:                      # a contiguous window of the raster order (generate_assistant.py:78-83)
                start = 0 if n == self.sample_quantity else \
                    int(np.random.choice(n - self.sample_quantity, 1, replace=False)[0])
                idx = np.arange(start, start + self.sample_quantity)
            idx = torch.as_tensor(idx, dtype=torch.long, device=device)
            coord = coord[idx]
            results["gt"] = results["gt"][idx]
            if "gt_unsharp" in results:
                results["gt_unsharp"] = results["gt_unsharp"][idx]
        results["coord"] = coord
        results["cell"] = _cell_for(coord, self.target_size)
        self._finish(results)
        return results

    def _finish(self, results):
        pass

    def __repr__(self):
        return (f"{self.__class__.__name__}sample_quantity={self.sample_quantity}, "
                f"scale={self.scale}, target_size={self.target_size}")


class GenerateCoordinateAndCell1(GenerateCoordinateAndCell):
    needs_unsharp = True

    def _finish(self, results):
        results["target_size"] = self.target_size            # generate_assistant.py:94


class GenerateCoordinateAndCell2:
    """Coordinates of the scale1 grid for a GT cropped for `scale` (generate_assistant.py:105-191)."""

    def __init__(self, sample_quantity=None, scale=None, scale1=None, target_size=None):
        self.sample_quantity = sample_quantity
        self.scale = scale
        self.scale1 = scale1
        self.target_size = target_size

    def __call__(self, results):
        device = results["gt"].device if "gt" in results else torch.device("cpu")
        if "gt" in results:
            _, h_hr, w_hr = results["gt"].shape
            results["gt"] = _flatten_rgb(results["gt"])
            h_lr, w_lr = h_hr / self.scale, w_hr / self.scale
            self.target_size = (round(h_lr * self.scale1), round(w_lr * self.scale1))
        else:
            assert self.target_size is not None
            assert len(self.target_size) >= 2
        coord = make_coord(self.target_size[-2:]).to(device)
        if self.sample_quantity is not None and "gt" in results:
            idx = torch.as_tensor(np.random.choice(len(coord), self.sample_quantity, replace=False),
                                  dtype=torch.long, device=device)
            coord = coord[idx]
            results["gt"] = results["gt"][idx]
        results["coord"] = coord
        results["cell"] = _cell_for(coord, self.target_size)
        return results

    def __repr__(self):
        return (f"{self.__class__.__name__}sample_quantity={self.sample_quantity}, "
                f"scale={self.scale}, target_size={self.target_size}")
