"""CPU oracle of the batched CVRP-TW construction environment (TEST INFRASTRUCTURE ONLY).

This file is a restatement, in plain torch-on-CPU fp32 arithmetic, of what the reference
environment computes on the hot path.  It exists to check the CUDA path; nothing under
``jampr_plus_b200/`` imports it, and the product never falls back to it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may execute it.

Pinning: the reference ships no golden vectors, so the pin is the reference itself,
run in the build container by ``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` replays
those traces through this oracle and demands bit-equal candidate sets, masks, step costs, tour buffers.

Reference lines restated (all under /root/reference/lib):
  load()        routing/env.py:344-418, utils/challenge_utils.py:23-35, utils/graph_utils.py:104-141
  reset()       routing/env.py:146-201;  pomo start  routing/env.py:1221-1297
  observe()     routing/env.py:848-907 (candidate sets), :932-1022 (feasibility mask), :1042-1073
  step()        routing/env.py:203-311, :1075-1131 (_update), :1035-1040, :1133-1170 (_return_all)
  graph_due()   routing/env.py:796-846 (which steps rebuild the tour graph)

State layout differs from the reference on purpose (it is the layout the CUDA kernels use):
static data is kept once per *instance* (the reference replicates it per sample, env.py:1189-1202),
and the tour graph is kept as predecessor/successor arrays (``pred``/``succ``) that are updated
per move instead of being re-derived from the tour buffers every step (env.py:804-835).
"""
import math
from typing import List, Optional, Sequence

import numpy as np
import torch

STATUS_NO_FEASIBLE = 1      # env.py:1020-1021
STATUS_TOUR_OVERFLOW = 2    # env.py:230-236
STATUS_FLEET_EXHAUSTED = 4  # env.py:254-258 (warning in the reference)


def dimacs_dist(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """floor(10 * ||100 (a-b)||) / 10, one correctly rounded fp32 operation per step (challenge_utils.py:35).

    The square root is taken with numpy (IEEE, correctly rounded) on purpose: torch.sqrt on the CPU goes
    through MKL's vsSqrt, which is off by one ulp on rare inputs and host dependent; behind the floor() that
    becomes a 0.1-unit jump.  The reference run on CUDA uses the correctly rounded sqrtf, as the kernels do.
    Traces recorded from the reference on the CPU are replayed with their own matrix (load(..., dist=))."""
    d = 100 * (a - b)
    s = (d * d).sum(-1)
    r = torch.from_numpy(np.sqrt(s.numpy()))
    return torch.floor(10 * r) / 10


class OracleEnv:
    def __init__(self, max_concurrent_vehicles: int = 3, k_nbh_frac: float = 0.25, pomo: bool = False,
                 pomo_nbh_frac: float = 0.25, pomo_single_start_node: bool = False, num_samples: int = 1,
                 tour_graph_update_step: int = 1):
        self.K = max_concurrent_vehicles
        self.k_nbh_frac = k_nbh_frac
        self.pomo = pomo
        self.pomo_nbh_frac = pomo_nbh_frac
        self.pomo_single = pomo_single_start_node
        self.P = num_samples
        self.upd = tour_graph_update_step

    # ------------------------------------------------------------------ static, per instance
    def load(self, instances: Sequence, dist: Optional[torch.Tensor] = None) -> None:
        f32 = torch.float32
        st = lambda key: torch.from_numpy(np.stack([np.asarray(x[key]) for x in instances])).to(f32)
        self.I = len(instances)
        self.coords = st("coords")
        self.demand = st("demands")
        self.tw = st("tw")
        self.service = st("service_time")
        self.horizon = st("org_service_horizon")
        self.N = N = int(instances[0].graph_size)
        kv = int(instances[0].max_vehicle_number)
        self.K_max = int(kv + np.floor(np.log(kv)))
        self.L = min(64, N)
        self.k = k = math.ceil(self.k_nbh_frac * N)
        self.dist = dimacs_dist(self.coords[:, :, None, :], self.coords[:, None, :, :]) / self.horizon[:, None, None]
        if dist is not None:      # replay of a recorded trace: the recorder's own transit matrix
            self.dist_formula = self.dist
            self.dist = dist.to(f32).clone()
        self.ttd = self.dist[:, :, 0].clone()
        # inference-time window fix-up (env.py:395-409)
        ret = self.tw[:, :, 1] + self.ttd + self.service[:, None]
        bad = ret > 1.0
        bad[:, 0] = False
        self.tw[:, :, 1][bad] = self.tw[:, :, 1][bad] - (ret[bad] - 1.0)
        if not bool(((self.tw[:, :, 1] - self.tw[:, :, 0])[bad] > 0.005).all()):
            raise AssertionError("time window collapses under the return-to-depot fix")
        # k-1 nearest customers (self included, ties -> lower index) + the depot as last entry
        c = self.coords[:, 1:, :]
        d = c[:, :, None, :] - c[:, None, :, :]
        d2 = (d * d).sum(-1)
        near = torch.sort(d2, dim=-1, stable=True).indices[:, :, : k - 1] + 1
        self.knn = torch.zeros(self.I, N, k, dtype=torch.long)
        self.knn[:, 1:, : k - 1] = near
        # all nodes by Euclidean distance to the depot, depot first (env.py:855-861)
        nrm = torch.norm(self.coords - self.coords[:, :1, :], p=2, dim=-1)
        self.order = torch.sort(nrm, dim=-1, stable=True).indices
        self.bs = self.I * self.P
        self.inst = torch.arange(self.I).repeat_interleave(self.P)

    def node_features(self) -> torch.Tensor:
        """(I, N, 7): coords, demand, tw, service time, time to depot (env.py:1047-1053)."""
        return torch.cat((self.coords, self.demand[:, :, None], self.tw,
                          self.service[:, None, None].expand(self.I, self.N, 1), self.ttd[:, :, None]), -1)

    # ------------------------------------------------------------------ episode state
    def reset(self, start_nodes: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None):
        bs, N, K = self.bs, self.N, self.K
        self.visited = torch.zeros(bs, N, dtype=torch.bool)
        self.finished = torch.zeros(bs, self.K_max, dtype=torch.bool)
        self.active = torch.zeros(bs, self.K_max, dtype=torch.bool)
        self.active[:, :K] = True
        self.plan = torch.zeros(bs, self.K_max, self.L, dtype=torch.int16)
        self.row = torch.arange(K).expand(bs, K).clone()
        self.nxt = torch.zeros(bs, K, dtype=torch.long)
        self.cur = torch.zeros(bs, K, dtype=torch.long)
        self.cap = torch.ones(bs, K)
        self.time = torch.zeros(bs, K)
        self.cttd = torch.zeros(bs, K)
        self.total = torch.zeros(bs)
        self.pred = torch.zeros(bs, N, dtype=torch.long)
        self.succ = torch.zeros(bs, N, dtype=torch.long)
        self.row_last = torch.zeros(bs, K, dtype=torch.long)
        self.status = torch.zeros(bs, dtype=torch.int32)
        self.step_no = 0
        self.done = False
        self.graph_fresh = True
        if self.pomo:
            self._pomo_start(start_nodes, generator)

    def pomo_candidates(self) -> torch.Tensor:
        """Candidate start nodes per instance: the depot neighbourhood without the depot, reduced to
        the floor((k-1) * frac) earliest window starts (env.py:1246-1252).  Ties in the window start
        are broken by neighbourhood position (the reference leaves it to an unstable sort)."""
        nb = self.order[:, 1: self.k]
        tws = self.tw[:, :, 0].gather(1, nb)
        srt = torch.sort(tws, dim=-1, stable=True).indices
        if self.pomo_single:
            return nb.gather(1, srt[:, : self.P])
        kk = math.floor((self.k - 1) * self.pomo_nbh_frac)
        return nb.gather(1, srt[:, :kk])

    def _pomo_start(self, start_nodes, generator):
        bs, K = self.bs, self.K
        if start_nodes is None:
            cand = self.pomo_candidates()
            if self.pomo_single:
                start_nodes = cand.reshape(-1, 1)
            else:
                pick = torch.multinomial(torch.ones(bs, cand.size(1)), K, replacement=False, generator=generator)
                start_nodes = cand[self.inst].gather(1, pick)
        start_nodes = start_nodes.long().view(bs, -1)
        cost = torch.zeros(bs)
        for t in range(start_nodes.size(1)):
            sel = torch.full((bs,), t, dtype=torch.long)
            cost = cost + self._move(sel, start_nodes[:, t])
        self.total += cost
        self.step_no += start_nodes.size(1)

    # ------------------------------------------------------------------ one transition (A.4 steps 1-5)
    def _move(self, sel: torch.Tensor, node: torch.Tensor) -> torch.Tensor:
        b = torch.arange(self.bs)
        ins = self.inst
        prev = self.cur[b, sel]
        self.cur[b, sel] = node
        self.cap[b, sel] = self.cap[b, sel] - self.demand[ins, node]
        delta = self.dist[ins, prev, node]
        arrival = self.time[b, sel] + delta
        nd = node != 0
        wait = (self.tw[ins, node, 0] - arrival).clamp(min=0) + self.service[ins]
        delta = torch.where(nd, delta + wait, delta)
        self.time[b, sel] = self.time[b, sel] + delta
        ttd_new = self.ttd[ins, node]
        cost = delta + (ttd_new - self.cttd[b, sel])
        self.cttd[b, sel] = ttd_new
        # tour buffer
        rows = self.row[b, sel]
        pos = self.nxt[b, sel]
        over = pos >= self.L
        self.status[over] |= STATUS_TOUR_OVERFLOW
        ok = ~over
        self.plan[b[ok], rows[ok], pos[ok]] = node[ok].to(torch.int16)
        self.nxt[b[nd], sel[nd]] = pos[nd] + 1
        # tour graph as predecessor / successor links along the plan row
        last = self.row_last[b, sel]
        bn, nn, ln = b[nd], node[nd], last[nd]
        self.pred[bn, nn] = ln
        has_prev = ln != 0
        self.succ[bn[has_prev], ln[has_prev]] = nn[has_prev]
        self.row_last[bn, sel[nd]] = nn
        self.visited[b, node] = True
        self.visited[:, 0] = False
        return cost

    def next_free(self) -> torch.Tensor:
        return torch.argmin((self.finished | self.active).to(torch.int), dim=-1)

    def step(self, action: torch.Tensor):
        """action (bs, 2) = (vehicle slot, node).  Returns (cost (bs,), done)."""
        assert not self.done
        b = torch.arange(self.bs)
        sel, node = action[:, 0].long(), action[:, 1].long()
        all_vis_before = self.visited[:, 1:].all(-1)
        cost = self._move(sel, node)
        ret = (node == 0) & ~all_vis_before
        if ret.any():
            nf = self.next_free()
            exhausted = ret & (nf == 0)
            self.status[exhausted] |= STATUS_FLEET_EXHAUSTED
            go = ret & (nf != 0)
            bb, ss, vv = b[go], sel[go], nf[go]
            old = self.row[bb, ss]
            self.cur[bb, ss] = 0
            self.cap[bb, ss] = 1.0
            self.time[bb, ss] = 0.0
            self.finished[bb, old] = True
            self.active[bb, old] = False
            self.active[bb, vv] = True
            self.nxt[bb, ss] = 0
            self.row[bb, ss] = vv
            self.row_last[bb, ss] = 0
        all_vis = self.visited[:, 1:].all(-1)
        done = bool(all_vis.all()) or self.step_no >= self.N + self.K_max + 1
        if done:
            en_route = self.cur != 0
            back = torch.where(en_route, self.dist[self.inst[:, None].expand_as(self.cur), self.cur, 0], torch.zeros(()))
            self.time = self.time + back
            self.cur = torch.zeros_like(self.cur)
            self.cttd = torch.zeros_like(self.cttd)
            cost = cost + back.sum(-1)
            unv = ~self.visited[:, 1:]
            if unv.any():
                pen = (self.ttd[self.inst][:, 1:] * 2)
                for r in unv.any(-1).nonzero().view(-1).tolist():
                    cost[r] = cost[r] + pen[r][unv[r]].sum()
        # tour-graph refresh rule for the observation that follows (env.py:801), pre-increment step
        self.graph_fresh = (self.step_no <= self.K + 1) or (self.step_no % self.upd == 0)
        self.total = self.total + cost
        self.step_no += 1
        self.done = done
        return cost, done

    # ------------------------------------------------------------------ solution import (LNS repair step)
    def import_sol(self, sol: List[List], cost: Optional[List] = None) -> None:
        """Restates RPEnv.import_sol (env.py:575-706) for POMO-style input (one solution per sample): complete
        tours [0, ..., 0] fill finished plan rows, other tours of length > 1 become active vehicles with
        recomputed capacity / clock (env.py:708-718, which charges a service at every list element, the depot
        entries of complete tours included), singletons are dropped; remaining slots get fresh vehicles in the
        first free rows; the total is the sum of the recomputed tour times when no cost is given; the step
        counter becomes max_b(visited + finished); the tour-graph refresh rule is evaluated with the step
        counter of the PREVIOUS episode (env.py:700 runs before :704)."""
        assert len(sol) == self.I and all(len(s) == self.P for s in sol), "one solution per sample expected"
        prev_step = self.step_no
        bs, N, K = self.bs, self.N, self.K
        self.visited = torch.zeros(bs, N, dtype=torch.bool)
        self.finished = torch.zeros(bs, self.K_max, dtype=torch.bool)
        self.active = torch.zeros(bs, self.K_max, dtype=torch.bool)
        self.plan = torch.zeros(bs, self.K_max, self.L, dtype=torch.int16)
        self.nxt = torch.zeros(bs, K, dtype=torch.long)
        self.cur = torch.zeros(bs, K, dtype=torch.long)
        self.cap = torch.ones(bs, K)
        self.time = torch.zeros(bs, K)
        self.cttd = torch.zeros(bs, K)
        self.row = torch.zeros(bs, K, dtype=torch.long)
        self.row_last = torch.zeros(bs, K, dtype=torch.long)
        self.status = torch.zeros(bs, dtype=torch.int32)
        totals = []

        def recompute(tour, i):
            tm = torch.zeros((), dtype=torch.float32)
            prev = 0
            for n in tour:
                tm = tm + self.dist[i, prev, n]
                tm = tm + ((self.tw[i, n, 0] - tm).clamp(min=0) + self.service[i])
                prev = n
            return float(tm)

        b = 0
        for i, inst_sol in enumerate(sol):
            for smp in inst_sol:
                num_partial, t_idx, costs = 0, 0, []
                for tour in smp:
                    if len(tour) <= 1:
                        continue
                    if t_idx >= self.K_max:
                        raise RuntimeError("Number of tours of provided solution is larger than max_num_vehicles!")
                    if tour[0] == 0 and tour[-1] == 0:
                        self.plan[b, t_idx, : len(tour) - 1] = torch.tensor(tour[1:], dtype=torch.int16)
                        self.finished[b, t_idx] = True
                        costs.append(recompute(tour, i))
                    else:
                        t = torch.tensor(tour, dtype=torch.long)
                        self.plan[b, t_idx, : len(tour)] = t.to(torch.int16)
                        self.cur[b, num_partial] = t[-1]
                        self.cap[b, num_partial] = 1.0 - self.demand[i].gather(0, t).sum()
                        self.cttd[b, num_partial] = self.ttd[i, t[-1]]
                        tm = recompute(tour, i)
                        costs.append(tm)
                        self.time[b, num_partial] = tm
                        self.nxt[b, num_partial] = len(tour)
                        self.active[b, t_idx] = True
                        self.row_last[b, num_partial] = t[-1]
                        num_partial += 1
                    t_idx += 1
                assert num_partial <= K
                for _ in range(num_partial, K):
                    nf = int(torch.argmin((self.finished[b] | self.active[b]).to(torch.int)))
                    self.active[b, nf] = True
                self.row[b] = self.active[b].nonzero().view(-1)          # env.py:697: ascending plan rows
                nz = self.plan[b][self.plan[b] > 0].long()
                self.visited[b, nz] = True
                totals.append(sum(costs))
                b += 1
        self.total = torch.tensor(totals, dtype=torch.float32) if cost is None else \
            torch.tensor(cost, dtype=torch.float32).view(-1)
        self.pred, self.succ = self.links_from_plan()
        self.graph_fresh = (prev_step <= K + 1) or (prev_step % self.upd == 0)
        self.step_no = int((self.visited.sum(-1) + self.finished.sum(-1)).max())
        self.done = False

    # ------------------------------------------------------------------ observation
    def observe(self):
        """Candidate sets nbh (bs, K, k) and infeasibility mask (bs, K, k)."""
        bs, K, k, N = self.bs, self.K, self.k, self.N
        ins = self.inst
        # depot vehicles: first k unvisited nodes in depot order, padded with the earliest visited ones
        order = self.order[ins]
        unv = ~self.visited.gather(1, order)
        n_unv = unv.sum(-1, keepdim=True)
        rank_unv = unv.cumsum(-1)
        rank_vis = (~unv).cumsum(-1)
        missing = (k - n_unv).clamp(min=0)
        pick = (unv & (rank_unv <= k)) | (~unv & (rank_vis <= missing))
        depot_nbh = order[pick].view(bs, k)
        nbh = self.knn[ins[:, None].expand(bs, K), self.cur]
        at_depot = self.cur == 0
        nbh = torch.where(at_depot[:, :, None], depot_nbh[:, None, :].expand(bs, K, k), nbh)
        bi = ins[:, None, None].expand(bs, K, k)
        vis = self.visited[torch.arange(bs)[:, None, None].expand(bs, K, k), nbh]
        ex_cap = self.cap[:, :, None] < self.demand[bi, nbh]
        arrival = self.time[:, :, None] + self.dist[bi, self.cur[:, :, None].expand(bs, K, k), nbh]
        ex_tw = arrival > self.tw[bi, nbh, 1]
        all_vis = self.visited[:, 1:].all(-1)
        mask_depot = ~all_vis & ((self.finished.sum(-1) < K) | (self.next_free() != 0))
        mask = vis.clone()
        mask[:, :, 0] |= at_depot & mask_depot[:, None]
        mask = mask | ex_cap | ex_tw
        dead = mask.all(-1).any(-1)
        self.status[dead] |= STATUS_NO_FEASIBLE
        return nbh, mask

    def tour_features(self) -> torch.Tensor:
        """(bs, K, 5) as env.py:1061-1068."""
        return torch.stack((self.cur.to(torch.float32), self.cap, self.time, self.cttd,
                            (-self.row + self.K_max) / self.K_max), -1)

    def active_rows(self) -> torch.Tensor:
        """(bs, K, L) tour buffers of the K active vehicles (env.py:1057-1059)."""
        return self.plan.gather(1, self.row[:, :, None].expand(self.bs, self.K, self.L)).long()

    def k_used(self) -> torch.Tensor:
        act = self.active.clone()
        act[self.active] = (self.cur != 0).view(-1)
        return (self.finished | act).sum(-1)

    # ------------------------------------------------------------------ helpers for tests
    def links_from_plan(self):
        """Re-derive pred/succ from the tour buffers the way env.py:804-835 builds tour edges."""
        pred = torch.zeros(self.bs, self.N, dtype=torch.long)
        succ = torch.zeros(self.bs, self.N, dtype=torch.long)
        sel = self.finished | self.active
        for b in range(self.bs):
            for r in sel[b].nonzero().view(-1).tolist():
                row = self.plan[b, r].long()
                src = torch.roll(row, 1)
                for s, d in zip(src.tolist(), row.tolist()):
                    if s != d:
                        if d != 0:
                            pred[b, d] = s
                        if s != 0:
                            succ[b, s] = d
        return pred, succ

    def true_cost(self) -> torch.Tensor:
        """Batch-independent cost of the tours in the buffers (logic of env.py:708-718 without the
        depot service quirk): travel + waiting + service along every tour, return leg included,
        plus 2 * time_to_depot for every unvisited customer."""
        out = torch.zeros(self.bs, dtype=torch.float64)
        for b in range(self.bs):
            i = int(self.inst[b])
            tot = 0.0
            for r in range(self.K_max):
                row = [int(v) for v in self.plan[b, r].tolist() if v > 0]
                if not row:
                    continue
                t, prev = 0.0, 0
                for n in row:
                    t += float(self.dist[i, prev, n])
                    t += max(float(self.tw[i, n, 0]) - t, 0.0) + float(self.service[i])
                    prev = n
                tot += t + float(self.dist[i, prev, 0])
            unv = (~self.visited[b, 1:]).nonzero().view(-1) + 1
            tot += float((2 * self.ttd[i, unv]).sum())
            out[b] = tot
        return out
