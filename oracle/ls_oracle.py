"""CPU twin of the device local search (TEST INFRASTRUCTURE ONLY): jampr_plus_b200/csrc/local_search.cu restated in
numpy fp32 scalar arithmetic — same move space (relocate, 2-opt*), same move ids, same evaluation by re-simulating the
touched tours with the environment's feasibility rules (reference lib/routing/env.py:981-990, 1081-1131), same
best-improvement rule (smallest (delta, move id) with delta < -1e-6).  The reference's own local search is OR-tools'
guided local search (lib/lns/gort.py:129-290), a third-party solver that is neither installed nor pinned: there is no
reference output to compare with, so the device search is pinned against this twin and against solution invariants."""
from typing import List, Tuple

import numpy as np

F = np.float32
EPS = F(1e-6)
INF = F(np.inf)


def sim(dist, tw, dem, srv, nodes: List[int]) -> np.float32:
    """Tour cost (duration incl. return leg) or +inf when a capacity / window test of the environment fails."""
    tm, cap, prev = F(0), F(1), 0
    if not nodes:
        return F(0)
    for n in nodes:
        d = dist[prev, n]
        if cap < dem[n]:
            return INF
        if F(tm + d) > tw[n, 1]:
            return INF
        cap = F(cap - dem[n])
        arrive = F(tm + d)
        tm = F(tm + F(d + F(max(F(tw[n, 0] - arrive), F(0)) + srv)))
        prev = n
    return F(tm + dist[prev, 0])


def local_search(plan: np.ndarray, dist, tw, dem, srv, max_iters: int) -> Tuple[np.ndarray, np.float32, int]:
    """plan (K_max, L) int16 -> (improved plan, cost of the served tours, moves applied)."""
    dist, tw, dem, srv = dist.astype(F), tw.astype(F), dem.astype(F), F(srv)
    R, L = plan.shape
    rows = []
    for r in range(R):
        l = 0
        while l < L and plan[r, l] != 0:
            l += 1
        rows.append([int(x) for x in plan[r, :l]])
    REL = R * L * R * (L + 1)
    moves = 0
    for _ in range(max_iters):
        cost = [sim(dist, tw, dem, srv, t) for t in rows]
        best = (-EPS, -1, None)

        def consider(delta, mid, new):
            nonlocal best
            if delta < best[0] or (delta == best[0] and mid < best[1]):
                best = (delta, mid, new)

        for r1 in range(R):
            l1 = len(rows[r1])
            if l1 == 0:
                continue
            empty_used = False
            for r2 in range(R):
                l2 = len(rows[r2])
                if l2 == 0:
                    if empty_used:
                        continue
                    empty_used = True
                if l2 >= L - 1 and r2 != r1:
                    continue
                for p1 in range(l1):
                    node = rows[r1][p1]
                    for p2 in range(l2 + 1):
                        mid = ((r1 * L + p1) * R + r2) * (L + 1) + p2
                        if r1 == r2:
                            if p2 == p1 or p2 == p1 + 1:
                                continue
                            t = rows[r1]
                            new = [x for q, x in enumerate(t[:p2]) if q != p1] + [node] + \
                                  [x for q, x in enumerate(t) if q >= p2 and q != p1]
                            consider(F(sim(dist, tw, dem, srv, new) - cost[r1]), mid, {r1: new})
                        else:
                            n1 = rows[r1][:p1] + rows[r1][p1 + 1:]
                            n2 = rows[r2][:p2] + [node] + rows[r2][p2:]
                            c1, c2 = sim(dist, tw, dem, srv, n1), sim(dist, tw, dem, srv, n2)
                            consider(F(F(c1 + c2) - F(cost[r1] + cost[r2])), mid, {r1: n1, r2: n2})
        for r1 in range(R):
            l1 = len(rows[r1])
            if l1 == 0:
                continue
            for r2 in range(r1 + 1, R):
                l2 = len(rows[r2])
                if l2 == 0:
                    continue
                for i in range(l1 + 1):
                    for j in range(l2 + 1):
                        if (i == l1 and j == l2) or (i == 0 and j == 0):
                            continue
                        if i + (l2 - j) > L - 1 or j + (l1 - i) > L - 1:
                            continue
                        n1 = rows[r1][:i] + rows[r2][j:]
                        n2 = rows[r2][:j] + rows[r1][i:]
                        c1, c2 = sim(dist, tw, dem, srv, n1), sim(dist, tw, dem, srv, n2)
                        mid = REL + ((r1 * (L + 1) + i) * R + r2) * (L + 1) + j
                        consider(F(F(c1 + c2) - F(cost[r1] + cost[r2])), mid, {r1: n1, r2: n2})
        if best[1] < 0:
            break
        for r, new in best[2].items():
            rows[r] = new
        moves += 1
    out = np.zeros_like(plan)
    for r, t in enumerate(rows):
        out[r, :len(t)] = t
    tot = F(0)
    for t in rows:
        tot = F(tot + sim(dist, tw, dem, srv, t))
    return out, tot, moves
