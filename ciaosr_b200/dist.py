"""Multi-GPU inference: one process per GPU, units of work sharded with no data-path
collective, and ONE all-gather of the final RGB (BASELINE.json north_star).

The path shards naturally: queries are independent given the feature map, batch items are
independent, and `clip_test` tiles are independent generator calls (ciaosr.py:233-245).
The reference itself only shards whole images across ranks (tools/test.py:123-146,
mmedit's multi_gpu_test); sharding *inside* a frame is this framework's addition.

Everything here works on any backend (`nccl` on GPUs, `gloo` in the CPU tests).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous, balanced [start, stop) of `n` units for `rank` (first n % world get one more)."""
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_round_robin(n, rank, world_size):
    """Indices rank, rank + world, ... : neighbouring tiles (similar cost) land on different ranks."""
    return list(range(rank, n, world_size))


def all_gather_padded(t, counts, group=None):
    """All-gather tensors whose dim 0 differs per rank.

    `counts[r]` = dim-0 length on rank r (known to every rank from the sharding rule, so no
    size exchange is needed).  One collective; returns the list of per-rank tensors.
    """
    rank, ws = world()
    if ws == 1:
        return [t]
    assert t.shape[0] == counts[rank], (t.shape, counts, rank)
    mx = max(counts)
    pad = t
    if t.shape[0] < mx:
        pad = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))], dim=0)
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return [o[:c] for o, c in zip(out, counts)]


def sharded_batch_forward(generator, lq, coord, cell, group=None):
    """Split the batch axis across ranks, run the generator on the local slice, all-gather
    the final RGB.  Every rank returns the full [B, Q, 3]."""
    rank, ws = world()
    b = lq.shape[0]
    counts = [shard_range(b, r, ws)[1] - shard_range(b, r, ws)[0] for r in range(ws)]
    s, e = shard_range(b, rank, ws)
    if e > s:
        local = generator(lq[s:e].contiguous(), coord[s:e].contiguous(), cell[s:e].contiguous(),
                          test_mode=True)
    else:
        local = lq.new_zeros((0, coord.shape[1], 3))
    return torch.cat(all_gather_padded(local, counts, group), dim=0)


def sharded_query_forward(generator, lq, coord, cell, eval_bsize=None, group=None):
    """Split the QUERY axis of one un-tiled call across ranks (BASELINE config 3, x6 / x8 on a whole frame):
    every rank runs the encoder and the cross-scale attention on the full LR input (they are replicated: a few
    per cent of the work at these scales), evaluates a contiguous band of the coordinate list, and one all-gather
    assembles [B, Q, 3].  Band boundaries are multiples of `eval_bsize`, because the reference reads tx, ty from
    the first cell of every eval_bsize-chunk (ciaosr_net.py:162-163 via :243): aligned bands see the same chunk
    starts as the unsharded call, so the result is identical."""
    rank, ws = world()
    q = coord.shape[1]
    unit = int(eval_bsize) if eval_bsize else 1
    n_units = -(-q // unit)
    bounds = [min(shard_range(n_units, r, ws)[0] * unit, q) for r in range(ws)] + [q]
    s, e = bounds[rank], bounds[rank + 1]
    if e > s:
        local = generator(lq, coord[:, s:e].contiguous(), cell[:, s:e].contiguous(), test_mode=True)
    else:
        local = lq.new_zeros((lq.shape[0], 0, 3))
    counts = [bounds[r + 1] - bounds[r] for r in range(ws)]
    parts = all_gather_padded(local.transpose(0, 1).contiguous(), counts, group)       # gather along the query axis
    return torch.cat(parts, dim=0).transpose(0, 1).contiguous()


def sharded_tile_predictions(origins, run_tile, tile_shape, like, group=None, run_many=None):
    """Deal `origins` (list of (y0, x0)) round-robin to the ranks, evaluate the local ones with
    `run_tile(y0, x0) -> [B, th*tw, 3]`, all-gather.  Returns predictions for ALL tiles in
    `origins` order, so that every rank can blend the full frame locally (the blend is the
    cheap E/W accumulate of ciaosr.py:253-255)."""
    rank, ws = world()
    n = len(origins)
    mine = shard_round_robin(n, rank, ws)
    counts = [len(shard_round_robin(n, r, ws)) for r in range(ws)]
    if mine:
        # run_many evaluates a list of tiles (possibly several per generator call), run_tile one at a time
        preds = run_many([origins[i] for i in mine]) if run_many is not None else [run_tile(*origins[i]) for i in mine]
        local = torch.stack(preds, dim=0)
    else:
        local = like.new_zeros((0,) + tuple(tile_shape))
    gathered = all_gather_padded(local, counts, group)
    out = [None] * n
    for r in range(ws):
        for j, i in enumerate(shard_round_robin(n, r, ws)):
            out[i] = gathered[r][j]
    return out
