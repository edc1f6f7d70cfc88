"""world_size-2 and -3 tests of the multi-GPU host logic on CPU (gloo): sharding rules, the single
padded all-gather, and that sharded evaluation reproduces the unsharded result."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ciaosr_b200 import dist as cd


def test_shard_rules():
    for n in [0, 1, 5, 16, 40, 220]:
        for ws in [1, 2, 3, 8]:
            cover = []
            for r in range(ws):
                s, e = cd.shard_range(n, r, ws)
                cover += list(range(s, e))
                assert 0 <= e - s <= -(-n // ws)
            assert cover == list(range(n))
            rr = sorted(i for r in range(ws) for i in cd.shard_round_robin(n, r, ws))
            assert rr == list(range(n))
    # BASELINE.json configs 4 and 5: 40 and 220 tiles over 8 GPUs
    assert [len(cd.shard_round_robin(40, r, 8)) for r in range(8)] == [5] * 8
    assert sorted({len(cd.shard_round_robin(220, r, 8)) for r in range(8)}) == [27, 28]


def _fake_generator(lq, coord, cell, test_mode=True):
    # any per-item, per-query function stands in for the head (which needs a GPU)
    b, q = coord.shape[:2]
    base = lq.mean(dim=(2, 3))[:, None, :]                      # [B,1,3]
    return base + coord.sum(-1, keepdim=True) * 0.5 + cell[..., :1] * torch.arange(3.0)


def _worker(rank, ws, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        g = torch.Generator().manual_seed(0)
        lq = torch.rand(5, 3, 6, 6, generator=g)
        coord = torch.rand(5, 11, 2, generator=g) * 2 - 1
        cell = torch.rand(5, 11, 2, generator=g)
        full = _fake_generator(lq, coord, cell)
        out = cd.sharded_batch_forward(_fake_generator, lq, coord, cell)
        assert out.shape == full.shape and torch.equal(out, full)
        # query-axis bands aligned to eval_bsize (chunk-start cells must be the same as unsharded)
        def chunky(lq_, coord_, cell_, test_mode=True, bs=4):
            outs = []
            for l in range(0, coord_.shape[1], bs):                  # like batched_predict: first cell of each chunk
                seed = cell_[:, l:l + 1, :1]
                outs.append(_fake_generator(lq_, coord_[:, l:l + bs], cell_[:, l:l + bs]) + seed)
            return torch.cat(outs, 1) if outs else lq_.new_zeros((lq_.shape[0], 0, 3))
        for bs in (4, 3, None):
            gen = (lambda a, b, c, test_mode=True, bs=bs: chunky(a, b, c, bs=bs)) if bs else _fake_generator
            ref = gen(lq, coord, cell)
            got = cd.sharded_query_forward(gen, lq, coord, cell, eval_bsize=bs)
            assert got.shape == ref.shape and torch.equal(got, ref), bs
        # uneven padded gather
        counts = [3, 1, 0, 2, 5][:ws]                    # includes a rank that contributes nothing
        t = torch.full((counts[rank], 2), float(rank))
        parts = cd.all_gather_padded(t, counts)
        assert [p.shape[0] for p in parts] == counts and float(parts[1][0, 0]) == 1.0
        # tiles
        origins = [(y, x) for y in (0, 4, 8) for x in (0, 4, 8, 12, 16)][:7]
        run = lambda y0, x0: torch.full((2, 4, 3), float(y0 * 100 + x0))
        preds = cd.sharded_tile_predictions(origins, run, (2, 4, 3), torch.zeros(1))
        assert [float(p[0, 0, 0]) for p in preds] == [float(y * 100 + x) for y, x in origins]
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws", [2, 3])
def test_multi_rank_gloo(tmp_path, ws):
    """ws = 3: 5 batch items, 11 queries and 7 tiles do not divide by the world size, and one rank gathers nothing."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(ws, port, str(tmp_path)), nprocs=ws, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(ws))
