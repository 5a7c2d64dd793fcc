"""Multi-GPU sharding of construction rollouts: one process per GPU, instances partitioned across ranks.

The hot path has no cross-rollout term: every rank loads a contiguous block of *instances* (all
samples of an instance stay on one rank, so the per-instance tables, the node-encoder output and the best-of-samples
reduction of ``training.py:80-94`` are rank-local) and runs the same fused device loop.  The only exchange steps are

  * ``gather_best``  one all_gather of (best cost f32, best sample i32, best tour plan i16[K_max x L]) per instance —
    ~3.6 KB per instance, latency-bound over NVLink (NCCL) — at the end of the episode;
  * ``global_done_step``  a 4-byte all_reduce(MAX) of the step at which each shard terminated: the reference's reported
    ``total_cost`` depends on the batch-global finishing step (env.py:281), so callers that compare
    ``total_cost`` across differently sharded runs need it; the batch-independent ``RPEnv.true_cost()`` does not.

Works with any ``torch.distributed`` backend: NCCL on the GPUs, gloo in the CPU tests (tests/test_sharding_gloo.py)."""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_list", "gather_best", "global_done_step", "lns_exchange"]


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``n_items`` owned by ``rank``: sizes differ by at most one, low ranks first."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_list(items: Sequence, rank: int, world: int) -> List:
    lo, hi = shard_bounds(len(items), rank, world)
    return list(items[lo:hi])


def gather_best(cost: torch.Tensor, idx: torch.Tensor, plan: torch.Tensor, n_total: int,
                group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather the per-instance best (cost (I_r,), sample index (I_r,), tour plan (I_r, K_max, L)) of every rank
    into global instance order.  ``n_total`` = number of instances over all ranks.  Ragged shards are padded to the
    largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return cost, idx, plan
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    assert cost.size(0) == sizes[rank][1] - sizes[rank][0], "local shard does not match shard_bounds()"
    cap = max(hi - lo for lo, hi in sizes)
    # one collective: every instance becomes one byte record [cost 4 B | idx 4 B | plan bytes]
    n_loc = cost.size(0)
    def rec_width(t):
        w = t.element_size()
        for s_ in t.shape[1:]:
            w *= int(s_)
        return w

    widths = [rec_width(t) for t in (cost, idx, plan)]
    rec = torch.zeros(cap, sum(widths), dtype=torch.uint8, device=cost.device)
    if n_loc > 0:
        rec[:n_loc] = torch.cat([t.contiguous().view(torch.uint8).reshape(n_loc, w)
                                 for t, w in zip((cost, idx, plan), widths)], 1)
    buf = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(buf, rec, group=group)
    allrec = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(buf, sizes)], 0)
    n = allrec.size(0)
    c0, c1 = widths[0], widths[0] + widths[1]
    g_cost = allrec[:, :c0].contiguous().view(cost.dtype).reshape((n,) + tuple(cost.shape[1:]))
    g_idx = allrec[:, c0:c1].contiguous().view(idx.dtype).reshape((n,) + tuple(idx.shape[1:]))
    g_plan = allrec[:, c1:].contiguous().view(plan.dtype).reshape((n,) + tuple(plan.shape[1:]))
    return g_cost, g_idx, g_plan


def global_done_step(local_done_step: int, device: torch.device, group: Optional[dist.ProcessGroup] = None) -> int:
    """Step at which the union batch would have terminated in the reference's lockstep loop (env.py:281)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(local_done_step)
    t = torch.tensor([int(local_done_step)], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def lns_exchange(total: torch.Tensor, plan: torch.Tensor, num_best: int, num_other: int, gen: torch.Generator,
                 group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Selection step of the sharded LNS (BASELINE config 5; lib/lns/lns.py:296-313 with the sub-problems split over
    the ranks): every rank contributes its ``num_best + num_other`` cheapest repaired solutions ``(total (P_r,), plan
    (P_r, K_max, L))``; one all_gather of costs and one of tour buffers (~3.6 KB per candidate over NVLink); every rank
    then keeps the globally cheapest ``num_best`` and ``num_other`` of the remaining candidates drawn with ``gen`` — a
    CPU generator seeded identically on all ranks, so all ranks continue from the same sources without a broadcast.
    Returns (costs (n_sel,), tour buffers (n_sel, K_max, L)).  Works on one rank (no collective) too."""
    n_sel = num_best + num_other
    order = total.sort(stable=True).indices[:n_sel]
    costs, plans = total[order].contiguous(), plan[order].contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        cl = [torch.empty_like(costs) for _ in range(world)]
        pb = plans.view(torch.uint8)            # int16 is not a collective dtype (NCCL / gloo): ship the bytes
        pl = [torch.empty_like(pb) for _ in range(world)]
        dist.all_gather(cl, costs, group=group)
        dist.all_gather(pl, pb, group=group)
        costs, plans = torch.cat(cl), torch.cat(pl).view(plans.dtype)
    o = costs.sort(stable=True).indices
    pick = o[:num_best]
    if num_other > 0:
        rest = o[num_best:]
        perm = torch.randperm(rest.numel(), generator=gen)[:num_other].to(rest.device)
        pick = torch.cat((pick, rest[perm]))
    return costs[pick], plans[pick].clone()
