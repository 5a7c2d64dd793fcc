"""LNS driver on the B200 construction / repair path: construct -> [select -> destroy -> repair]*.

Mirrors the search loop of reference ``lib/lns/lns.py:236-372`` for the part this repo accelerates: model
construction (``_construct``, :236-275) on ``RPEnv`` / ``Policy`` / the fused device rollout, selection of
``num_best + num_other`` samples (``RPEnv.export_sol``), the incumbent bookkeeping (``_select_best`` / ``_add_best``,
:175-222), the bootstrap of the selected solutions back to ``num_samples`` and their destruction (``_destruct``,
:296-313 with ``lib/lns/destructor.py:64-205``: random_from_route with random_nodes / random_cut / const_cut and
random / smallest complete-route removal), ``RPEnv.import_sol`` and the repair rollout.

NOT here: the OR-tools guided local search (``lib/lns/gort.py``; OR-tools is neither installed nor pinned by the
reference), the ``similarity`` / ``waiting_time`` complete-route removal modes of the destructor (they need the
local search's infos) and the DIMACS controller protocol.  Costs are reported two ways: the environment's own total (what the reference's loop compares, SURVEY.md
§0.6 / A.6 explains its quirks) and the batch-independent tour cost in original units (``true_cost * horizon``).
Host-side list handling stays in numpy / Python like the reference's; the model work runs through the CUDA library
(no CPU fallback)."""
import time
from copy import deepcopy
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .driver import rollout
from .routing.formats import RPInstance

__all__ = ["read_solomon", "Destructor", "LNS"]


def read_solomon(path: str) -> RPInstance:
    """Solomon / Gehring-Homberger text file -> normalised RPInstance (io_utils.py:24-128): coordinates / 100,
    demand / capacity, windows and service time / horizon (= depot due date)."""
    rows, kv, cap = [], None, None
    with open(path, "rt") as f:
        for i, line in enumerate(f, start=1):
            v = line.split()
            if i == 5:
                kv, cap = int(v[0]), float(v[1])
            elif i >= 10 and len(v) == 7:
                rows.append([float(x) for x in v])
    a = np.asarray(rows, dtype=np.float64)
    horizon = a[0, 5]
    coords = a[:, 1:3] / 100.0
    demand = a[:, 3] / cap
    tw = a[:, 4:6] / horizon
    service = a[:, 6] / horizon
    assert np.all(service[1:] == service[1])
    has_tw = a[1:, 4] != 0
    return RPInstance(coords=coords, demands=demand, tw=tw, service_time=service[1], graph_size=coords.shape[0],
                      org_service_horizon=horizon, max_vehicle_number=kv, vehicle_capacity=1.0, service_horizon=1.0,
                      depot_idx=[0], type="1" if cap < 500 else "2", tw_frac=str(has_tw.sum() / has_tw.size))


class Destructor:
    """Restatement of ``lib/lns/destructor.py`` (random_from_route family).  Solutions are lists of tours; complete
    tours are emitted as ``[0, ..., 0]``, destroyed ones as open customer lists (the convention import_sol reads,
    env.py:578-581)."""
    RM_P_MODES = ("random_nodes", "random_cut", "const_cut")

    def __init__(self, num_partial: int, rm_partial_mode: str = "random", rm_complete_mode: str = "random",
                 frac_rm_nodes_from_partial: float = 0.5, frac_rm_complete_routes: float = 0.15, seed: int = 1):
        assert rm_partial_mode in self.RM_P_MODES + ("random",)
        assert rm_complete_mode in ("random", "smallest")
        assert 0 <= frac_rm_nodes_from_partial < 1 and 0 <= frac_rm_complete_routes < 1
        self.num_partial = num_partial
        self.rm_partial_mode = rm_partial_mode
        self.rm_complete_mode = rm_complete_mode
        self.frac_p = frac_rm_nodes_from_partial
        self.frac_c = frac_rm_complete_routes
        self.rnd = np.random.default_rng(seed)

    def _rnd_int_1(self, n: int) -> int:          # destructor.py:58-63: 1 <= x < n-1, or 1
        return 1 if n <= 2 else int(self.rnd.integers(1, n - 1))

    def _partial(self, tour: List[int]) -> List[int]:
        if tour[0] == 0 and tour[-1] == 0:
            tour = tour[1:-1]
        mode = self.rm_partial_mode
        if mode == "random":
            mode = str(self.rnd.choice(self.RM_P_MODES))
        n = len(tour)
        if mode == "random_nodes":               # destructor.py:141-149
            n_rm = int(max(round(self.frac_p * n), 1)) if self.frac_p > 0 else self._rnd_int_1(n)
            rm = set(self.rnd.choice(np.arange(n), size=min(n_rm, n), replace=False).tolist())
            return [e for i, e in enumerate(tour) if i not in rm]
        if mode == "random_cut":                 # :151-155
            return tour[: self._rnd_int_1(n)]
        assert self.frac_p > 0                   # const_cut :157-161
        return tour[: -int(max(round(self.frac_p * n), 1))]

    def _rm_complete(self, tours: List[List[int]]) -> List[List[int]]:
        if self.frac_c <= 0 or not tours:
            return tours
        n_c = len(tours)
        n_rm = int(max(round(self.frac_c * n_c), 1))
        if self.rm_complete_mode == "random":
            rm = set(self.rnd.choice(np.arange(n_c), size=n_rm, replace=False).tolist())
        else:
            rm = set(np.argsort(np.array([len(t) for t in tours]))[:n_rm].tolist())
        return [t for i, t in enumerate(tours) if i not in rm]

    def random_from_route(self, sol: List[List[int]]) -> List[List[int]]:     # destructor.py:86-121
        lens = np.array([len(t) for t in sol])
        closed = sol[0][0] == 0 and sol[0][-1] == 0
        max_idx = 3 if closed else 1
        cand = np.flatnonzero(lens > max_idx)
        pick = set(self.rnd.choice(cand, min(self.num_partial, cand.size), replace=False).tolist()) if cand.size else set()
        c_tours, p_tours = [], []
        for i, tour in enumerate(sol):
            if i in pick:
                p_tours.append(self._partial(list(tour)))
            elif len(tour) > max_idx:
                c_tours.append(list(tour) if closed else [0] + list(tour) + [0])
        return self._rm_complete(c_tours) + p_tours

    def destruct(self, solutions: Sequence[List[List[int]]]) -> List[List[List[int]]]:
        return [self.random_from_route(s) for s in solutions]


class LNS:
    """``LNS(instance, policy, ...).search(time_limit=..)`` — the reference's loop without local search."""

    def __init__(self, instance: RPInstance, policy, env, num_best: int = 2, num_other: int = 3,
                 add_best: bool = True, destructor: Optional[Destructor] = None, seed: int = 1234,
                 use_graph: bool = True, selection_mode: str = "similarity", ls_step: int = 0, ls_iters: int = 200):
        assert env.inference and env.pomo, "LNS construction uses POMO inference (search/default.yaml: pomo_inference)"
        P = env.num_samples
        assert P % (num_best + num_other) == 0, "num_samples must be a multiple of num_best + num_other (lns.py:307)"
        self.inst, self.policy, self.env = instance, policy, env
        self.num_best, self.num_other, self.add_best = num_best, num_other, add_best
        self.ds = destructor or Destructor(num_partial=env.max_concurrent_vehicles, seed=seed)
        self.use_graph = use_graph
        # the reference runs OR-tools' guided local search every ls_step iterations (search/default.yaml ls_step,
        # lns.py `_local_search`); here: the device local search (RPEnv.local_search), 0 = off
        self.ls_step, self.ls_iters = int(ls_step), int(ls_iters)
        self.selection_mode = selection_mode          # config_lns/search/default.yaml: "similarity"
        self.horizon = float(instance.org_service_horizon)
        self.best_solution, self.best_cost, self.best_true_cost, self.best_step = None, float("inf"), float("inf"), -1
        self.step = 0
        self.seed = seed
        self.history: List[Tuple[float, float, float]] = []      # (seconds, incumbent env cost, incumbent true cost)
        env.seed(seed)

    # ------------------------------------------------------------------ one construction / repair (lns.py:236-275)
    @torch.no_grad()
    def _construct(self, partial: Optional[List[List[List[int]]]] = None):
        env, pol = self.env, self.policy
        if partial is None:
            pol.set_decode_type("greedy")
            pol.reset_static()
            env.load_data([self.inst])
            env.reset()
        else:
            env.import_sol([partial])
        rollout(pol, env, graph=self.use_graph and partial is None)
        if self.ls_step > 0:
            if self.step % self.ls_step == 0:
                env.local_search(self.ls_iters)
            else:
                env.state["total"].copy_(env.true_cost())
        true = env.true_cost() * self.horizon
        sols, costs = env.export_sol(self.num_best, self.num_other, mode=self.selection_mode)
        # recover which samples were exported to attach their batch-independent cost
        total = env.state["total"].cpu().numpy()
        tc = true.cpu().numpy()
        out_true = []
        for c in costs[0]:
            j = int(np.flatnonzero(total == np.float32(c))[0])
            out_true.append(float(tc[j]))
        return sols[0], (np.asarray(costs[0]) * self.horizon).tolist(), out_true

    def _select_best(self, sols, costs, true_costs):
        j = int(np.argmin(costs))
        if self.best_solution is None or costs[j] < self.best_cost:
            self.best_solution, self.best_cost, self.best_true_cost = deepcopy(sols[j]), costs[j], true_costs[j]
            self.best_step = self.step

    def _add_best(self, sols, costs, true_costs):                 # lns.py:209-222
        if self.best_cost not in costs:
            w = int(np.argmax(costs))
            sols[w], costs[w], true_costs[w] = deepcopy(self.best_solution), self.best_cost, self.best_true_cost
        return sols, costs, true_costs

    # ------------------------------------------------------------------ the same loop with every solution on the device
    @torch.no_grad()
    def search_device(self, time_limit: Optional[float] = None, num_steps: Optional[int] = None,
                      group=None) -> Tuple[torch.Tensor, float, float]:
        """construct -> [select -> add incumbent -> bootstrap -> destroy -> import -> repair]* with tour buffers that
        never leave the device (``RPEnv.destruct``: selection mode "random", the destructor's random_from_route family).
        Per iteration the host reads two scalars (the new step counter, the incumbent cost for the history).

        ``group`` / an initialised default process group with more than one rank: BASELINE config 5 — instance and
        weights replicated, every rank mutilates and repairs its own ``num_samples`` sub-problems with its own random
        streams; after each repair the ranks all_gather their ``num_best + num_other`` cheapest candidates (cost + tour
        buffer, ~3.6 KB each, NCCL over NVLink; gloo in the CPU tests), every rank keeps the globally cheapest
        ``num_best`` plus ``num_other`` others drawn with a shared seed as the sources of its next iteration.
        Returns (incumbent tour buffer (K_max, L) int16, incumbent env cost, incumbent true cost), original units."""
        import torch.distributed as dist
        assert time_limit is not None or num_steps is not None
        env, pol, ds = self.env, self.policy, self.ds
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        n_sel = self.num_best + self.num_other
        t0 = time.time()
        pol.set_decode_type("greedy")
        pol.reset_static()
        env.load_data([self.inst])
        env.reset()
        rollout(pol, env, graph=self.use_graph)
        if self.ls_step > 0:
            env.local_search(self.ls_iters)
        kw = dict(num_best=self.num_best, num_other=self.num_other, num_partial=ds.num_partial,
                  rm_partial_mode=ds.rm_partial_mode, rm_complete_mode=ds.rm_complete_mode,
                  frac_rm_nodes_from_partial=ds.frac_p, frac_rm_complete_routes=ds.frac_c,
                  seed=self.seed * 1000003 + rank)
        inc_cost = torch.full((1,), float("inf"), device=env.device)
        inc_true = torch.full((1,), float("inf"), device=env.device)
        inc_plan = torch.zeros(1, env.state.K_max, env.state.L, dtype=torch.int16, device=env.device)
        gen = torch.Generator(device="cpu").manual_seed(self.seed)       # shared by all ranks: same "others"

        def incumbent_update():
            cost, idx, plan = env.gather_best()
            true = env.true_cost().view(1, -1).gather(1, idx.long().view(1, 1)).view(1) * self.horizon
            better = cost < inc_cost
            inc_cost.copy_(torch.where(better, cost, inc_cost))
            inc_true.copy_(torch.where(better, true, inc_true))
            inc_plan.copy_(torch.where(better.view(1, 1, 1), plan, inc_plan))

        def global_sources():
            """Sharded selection (sharding.lns_exchange) + the incumbent rule of lns.py:209-222."""
            from .sharding import lns_exchange
            costs, src = lns_exchange(env.state["total"], env.state["plan"], self.num_best, self.num_other, gen, group)
            if self.add_best and not bool((costs == inc_cost).any()):
                src[costs.argmax()] = inc_plan[0]
            return src.unsqueeze(0)

        def sync_incumbent():
            if world == 1:
                return
            c = [torch.empty_like(inc_cost) for _ in range(world)]
            dist.all_gather(c, inc_cost, group=group)
            best_rank = int(torch.cat(c).argmin())
            src = dist.get_global_rank(group, best_rank) if group is not None else best_rank
            for t in (inc_cost, inc_true, inc_plan.view(torch.uint8)):      # int16 is not a collective dtype
                dist.broadcast(t, src=src, group=group)

        incumbent_update()
        sync_incumbent()
        hz = self.horizon
        self.history.append((time.time() - t0, float(inc_cost) * hz, float(inc_true)))
        slowest = 0.0
        while True:
            if num_steps is not None and self.step >= num_steps:
                break
            if time_limit is not None:
                stop = (time.time() - t0) >= max(time_limit - slowest, 0.0)
                if world > 1:        # every rank must leave the loop in the same iteration
                    flag = torch.tensor([int(stop)], device=env.device)
                    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
                    stop = bool(flag.item())
                if stop:
                    break
            ts = time.time()
            self.step += 1
            if world > 1:
                env.destruct(sources=global_sources(), iteration=self.step, **kw)
            else:
                env.destruct(iteration=self.step, incumbent_plan=inc_plan if self.add_best else None,
                             incumbent_cost=inc_cost, **kw)
            rollout(pol, env, graph=self.use_graph)
            if self.ls_step > 0 and self.step % self.ls_step == 0:
                env.local_search(self.ls_iters)
            elif self.ls_step > 0:
                env.state["total"].copy_(env.true_cost())      # one cost scale for improved and untouched solutions
            incumbent_update()
            sync_incumbent()
            slowest = max(slowest, time.time() - ts)
            self.history.append((time.time() - t0, float(inc_cost) * hz, float(inc_true)))
        self.best_cost, self.best_true_cost = float(inc_cost) * hz, float(inc_true)
        self.best_plan = inc_plan[0].clone()
        return self.best_plan, self.best_cost, self.best_true_cost

    def search(self, time_limit: Optional[float] = None, num_steps: Optional[int] = None):
        assert time_limit is not None or num_steps is not None
        t0 = time.time()
        P = self.env.num_samples
        sols, costs, tcs = self._construct()
        self._select_best(sols, costs, tcs)
        self.history.append((time.time() - t0, self.best_cost, self.best_true_cost))
        slowest = 0.0
        while True:
            if num_steps is not None and self.step >= num_steps:
                break
            if time_limit is not None and (time.time() - t0) >= max(time_limit - slowest, 0.0):
                break
            ts = time.time()
            if self.add_best:
                sols, costs, tcs = self._add_best(sols, costs, tcs)
            boot = [deepcopy(s) for s in sols] * (P // len(sols))              # lns.py:304-310
            partial = self.ds.destruct(boot)
            self.step += 1
            sols, costs, tcs = self._construct(partial)
            self._select_best(sols, costs, tcs)
            slowest = max(slowest, time.time() - ts)
            self.history.append((time.time() - t0, self.best_cost, self.best_true_cost))
        return self.best_solution, self.best_cost, self.best_true_cost
