"""World-size-2 (and 3, ragged) gloo runs of the multi-GPU host logic (jampr_plus_b200/sharding.py) on the CPU:
partitioning, the best-tour all_gather in global instance order, the finishing-step all_reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jampr_plus_b200.sharding import gather_best, global_done_step, shard_bounds, shard_list


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert shard_list(list("abcdefg"), 1, 3) == ["d", "e"]
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        cost = torch.rand(n_total, generator=g)
        idx = torch.randint(0, 16, (n_total,), generator=g, dtype=torch.int32)
        plan = torch.randint(0, 101, (n_total, 5, 8), generator=g, dtype=torch.int16)
        lo, hi = shard_bounds(n_total, rank, world)
        c, i, p = gather_best(cost[lo:hi], idx[lo:hi], plan[lo:hi], n_total)
        assert torch.equal(c, cost) and torch.equal(i, idx) and torch.equal(p, plan)
        assert global_done_step(100 + 7 * rank, torch.device("cpu")) == 100 + 7 * (world - 1)
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 10), (3, 10), (2, 1)])
def test_gather_best_gloo(tmp_path, world, n_total):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(tmp_path / f"ok{r}.npy")


def _lns_worker(rank, world, port, out_dir):
    from jampr_plus_b200.sharding import lns_exchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, R, L, nb, no = 10, 4, 6, 2, 3
        g = torch.Generator().manual_seed(100 + rank)                   # every rank repaired different sub-problems
        total = torch.rand(P, generator=g)
        plan = torch.randint(0, 50, (P, R, L), generator=g, dtype=torch.int16)
        shared = torch.Generator().manual_seed(5)                       # same on all ranks
        costs, src = lns_exchange(total, plan, nb, no, shared)
        # reference: the union of all ranks' candidates, computed locally from the known seeds
        allc, allp = [], []
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            t = torch.rand(P, generator=gr)
            p = torch.randint(0, 50, (P, R, L), generator=gr, dtype=torch.int16)
            o = t.sort(stable=True).indices[: nb + no]
            allc.append(t[o])
            allp.append(p[o])
        allc, allp = torch.cat(allc), torch.cat(allp)
        o = allc.sort(stable=True).indices
        assert torch.equal(costs[:nb], allc[o[:nb]]) and torch.equal(src[:nb], allp[o[:nb]])
        rest = o[nb:]
        perm = torch.randperm(rest.numel(), generator=torch.Generator().manual_seed(5))[:no]
        assert torch.equal(costs[nb:], allc[rest[perm]]) and torch.equal(src[nb:], allp[rest[perm]])
        # every rank holds the same selection
        mine = torch.cat((costs, src.float().view(-1)))
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        assert all(torch.equal(x, mine) for x in got)
        np.save(os.path.join(out_dir, f"lns{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_lns_exchange_gloo(tmp_path, world):
    """Selection step of the sharded LNS (config 5): candidates of all ranks, globally best + shared random others."""
    port = _free_port()
    mp.spawn(_lns_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(tmp_path / f"lns{r}.npy")
