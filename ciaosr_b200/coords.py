"""Coordinate helpers of the boundary (mmedit 0.11 ``make_coord`` is not vendored by
the reference; call sites: ciaosr_net.py:148, ciaosr.py:240, generate_assistant.py:70)."""
import torch


def make_coord(shape, ranges=None, flatten=True):
    """Pixel-centre coordinates of a grid, (y, x) order, in [-1, 1] by default.

    Per dim n: r = (v1 - v0) / (2n); seq = v0 + r + 2r * arange(n) (fp32).
    """
    seqs = []
    for i, n in enumerate(shape):
        v0, v1 = (-1, 1) if ranges is None else ranges[i]
        r = (v1 - v0) / (2 * n)
        seqs.append(v0 + r + (2 * r) * torch.arange(n).float())
    ret = torch.stack(torch.meshgrid(*seqs, indexing="ij"), dim=-1)
    if flatten:
        ret = ret.view(-1, ret.shape[-1])
    return ret


def make_cell(target_hw, n):
    """cell = (2/Ht, 2/Wt) per query, as built at ciaosr.py:241-243."""
    cell = torch.ones(n, 2)
    cell[:, 0] *= 2 / target_hw[0]
    cell[:, 1] *= 2 / target_hw[1]
    return cell
