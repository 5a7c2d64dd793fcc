"""CPU twin of the device destruction operator (TEST INFRASTRUCTURE ONLY).

`destroy_plan` restates jampr_plus_b200/csrc/lns.cu `lns_destroy_kernel` — itself a restatement of the reference's
lib/lns/destructor.py:86-205 `random_from_route` (random_nodes :141-149, random_cut :151-155, const_cut :157-161,
random / smallest complete-route removal :186-205) on tour buffers — with the same random numbers
(oracle/rng.py twin of jampr_uniform, one stream per rollout and LNS iteration, draws in the kernel's order).
The reference draws from numpy's Generator, whose stream cannot be matched on a device: parity with the reference is
by operator semantics (tests/test_lns_device.py checks them), parity between device and twin is bit-exact."""
from typing import List, Tuple

import numpy as np

from .rng import uniform_scalar

LNS_DRAWS = 8192


class _Draws:
    def __init__(self, seed, b, iteration):
        self.seed, self.b, self.base, self.c = seed, b, iteration * LNS_DRAWS, 0

    def next(self) -> np.float32:
        u = uniform_scalar(self.seed, self.b, self.base + self.c)
        self.c += 1
        return np.float32(u)


def _rnd_int_1(r: _Draws, n: int) -> int:
    if n <= 2:
        return 1
    u = r.next()
    return 1 + min(int(np.float32(u * np.float32(n - 2))), n - 3)


def _smallest(keys: List[np.float32], n_rm: int) -> List[bool]:
    rm = [False] * len(keys)
    for _ in range(min(n_rm, len(keys))):
        best = -1
        for i, k in enumerate(keys):
            if not rm[i] and (best < 0 or k < keys[best]):
                best = i
        rm[best] = True
    return rm


def _rint(x: np.float32) -> int:
    return int(np.rint(np.float32(x)))


def destroy_plan(plan: np.ndarray, b: int, seed: int, iteration: int, num_partial: int, rm_partial_mode: int,
                 rm_complete_mode: int, frac_partial: float, frac_complete: float) -> Tuple[List[List[int]], List[int]]:
    """plan (K_max, L) int16 tour buffer of ONE solution -> (tours, kinds) in the dense order of the kernel
    (kept complete tours, kind 1, then mutilated tours, kind 2)."""
    fp, fc = np.float32(frac_partial), np.float32(frac_complete)
    rows = []
    for r in range(plan.shape[0]):
        row = plan[r]
        l = 0
        while l < row.size and row[l] != 0:
            l += 1
        rows.append([int(x) for x in row[:l]])
    rng = _Draws(seed, b, iteration)
    role = [1 if len(t) > 1 else 0 for t in rows]
    cand = [r for r, ro in enumerate(role) if ro]
    keys = [rng.next() for _ in cand]
    for i, flag in enumerate(_smallest(keys, min(num_partial, len(cand)))):
        if flag:
            role[cand[i]] = 2
    partial = []
    for r, tour in enumerate(rows):
        if role[r] != 2:
            continue
        n = len(tour)
        mode = rm_partial_mode
        if mode == 3:
            mode = min(int(np.float32(rng.next() * np.float32(3.0))), 2)
        if mode == 0:
            n_rm = max(_rint(fp * np.float32(n)), 1) if fp > 0 else _rnd_int_1(rng, n)
            n_rm = min(n_rm, n)
            nk = [rng.next() for _ in range(n)]
            rm = _smallest(nk, n_rm)
            partial.append([e for e, f in zip(tour, rm) if not f])
        elif mode == 1:
            partial.append(tour[: min(_rnd_int_1(rng, n), n)])
        else:
            cut = max(_rint(fp * np.float32(n)), 1)
            partial.append(tour[: max(n - cut, 0)])
    comp = [r for r, ro in enumerate(role) if ro == 1]
    if fc > 0 and comp:
        n_rm = max(_rint(fc * np.float32(len(comp))), 1)
        keys = [rng.next() if rm_complete_mode == 0 else np.float32(len(rows[r])) for r in comp]
        rm = _smallest(keys, n_rm)
        comp = [r for r, f in zip(comp, rm) if not f]
    tours = [rows[r] for r in comp] + partial
    kinds = [1] * len(comp) + [2] * len(partial)
    return tours, kinds
