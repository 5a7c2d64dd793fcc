"""The destruction operator on tour buffers (oracle/destroy_oracle.py, the CPU twin of csrc/lns.cu) against the
semantics of the reference's lib/lns/destructor.py:86-205, run unmodified where a copy of the reference is available:
both are random, with different generators, so the comparison is on what the operators guarantee — which tours may be
touched, how many customers / tours go, what is emitted — over many seeds, plus the distribution of the cut points."""
import importlib.util
import os

import numpy as np
import pytest

from oracle.destroy_oracle import destroy_plan
from oracle import ref_runner

MODES = {"random_nodes": 0, "random_cut": 1, "const_cut": 2, "random": 3}


def _plan(rng, R=12, L=16, n_tours=7):
    """A tour buffer with tours of assorted lengths (incl. singletons) over distinct customers."""
    plan = np.zeros((R, L), dtype=np.int16)
    perm = rng.permutation(np.arange(1, 100))
    pos = 0
    lens = [1, 1, 2, 3, 5, 8, 12][:n_tours]
    rng.shuffle(lens)
    for r, l in enumerate(lens):
        plan[r, :l] = perm[pos:pos + l]
        pos += l
    return plan, lens


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("cmode", [0, 1])
def test_operator_guarantees(mode, cmode):
    rng = np.random.default_rng(3)
    for trial in range(60):
        plan, lens = _plan(rng)
        fp, fc, npart = 0.5, 0.3, 3
        tours, kinds = destroy_plan(plan, b=trial, seed=99, iteration=trial % 5, num_partial=npart,
                                    rm_partial_mode=MODES[mode], rm_complete_mode=cmode, frac_partial=fp, frac_complete=fc)
        src = {tuple(int(x) for x in plan[r, :l]) for r, l in enumerate(lens)}
        cand = [t for t in src if len(t) > 1]
        comp = [t for t, k in zip(tours, kinds) if k == 1]
        part = [t for t, k in zip(tours, kinds) if k == 2]
        assert kinds == sorted(kinds), "complete tours first, then the mutilated ones (destructor.py:121)"
        assert len(part) == min(npart, len(cand))
        # complete tours are untouched source tours with more than one customer; singletons never survive
        assert all(tuple(t) in src and len(t) > 1 for t in comp)
        n_c = len(cand) - len(part)
        assert len(comp) == n_c - max(int(round(fc * n_c)), 1)
        if cmode == 1 and comp:          # smallest tours went first
            removed = [t for t in cand if list(t) not in comp and not any(_is_sub(p, t) for p in part)]
            assert max(len(t) for t in removed) <= min(len(t) for t in comp) or len(removed) == 0
        # every mutilated tour is an order-preserving strict sub-sequence of a distinct candidate
        used = set()
        for p in part:
            owners = [t for t in cand if t not in used and list(t) not in comp and _is_sub(p, t) and len(p) < len(t)]
            assert owners, (p, cand)
            used.add(owners[0])
            n = len(owners[0])
            if mode == "const_cut":
                assert len(p) == max(n - max(int(round(fp * n)), 1), 0) and list(owners[0][:len(p)]) == p
            if mode == "random_cut":
                assert 1 <= len(p) <= max(n - 2, 1) and list(owners[0][:len(p)]) == p
            if mode == "random_nodes":
                assert len(p) == n - min(max(int(round(fp * n)), 1), n)


def _is_sub(p, t):
    it = iter(t)
    return all(any(x == y for y in it) for x in p)


def test_cut_points_are_uniform():
    """random_cut draws the cut uniformly from [1, n - 2] (destructor.py:58-63, 151-155)."""
    plan = np.zeros((4, 16), dtype=np.int16)
    plan[0, :12] = np.arange(1, 13)
    counts = np.zeros(13, dtype=int)
    for b in range(4000):
        tours, kinds = destroy_plan(plan, b=b, seed=7, iteration=0, num_partial=1, rm_partial_mode=1,
                                    rm_complete_mode=0, frac_partial=0.0, frac_complete=0.0)
        counts[len(tours[-1])] += 1
    assert counts[0] == 0 and counts[11:].sum() == 0
    exp = 4000 / 10
    assert np.abs(counts[1:11] - exp).max() < 5 * np.sqrt(exp)


@pytest.mark.skipif(ref_runner.reference_root() is None, reason="no copy of the reference available")
def test_same_guarantees_as_the_reference_destructor():
    """The unmodified reference operator on the same solutions: same number of complete / mutilated tours and the same
    multiset of possible mutilated lengths for the deterministic const_cut mode."""
    path = os.path.join(ref_runner.reference_root(), "lib", "lns", "destructor.py")
    ref_runner.load_reference()
    spec = importlib.util.spec_from_file_location("ref_destructor", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(11)
    for trial in range(25):
        plan, lens = _plan(rng)
        sol = [[int(x) for x in plan[r, :l]] for r, l in enumerate(lens)]
        ds = mod.Destructor(num_partial=3, destruction_mode="random_from_route", rm_partial_mode="const_cut",
                            rm_complete_mode="smallest", frac_rm_nodes_from_partial=0.4, frac_rm_complete_routes=0.25,
                            seed=trial)
        out = ds.destruct([sol])[0]
        r_comp = [t for t in out if t[0] == 0 and t[-1] == 0]
        r_part = [t for t in out if not (t[0] == 0 and t[-1] == 0)] if all(len(t) for t in out) else None
        tours, kinds = destroy_plan(plan, b=trial, seed=5, iteration=0, num_partial=3, rm_partial_mode=2,
                                    rm_complete_mode=1, frac_partial=0.4, frac_complete=0.25)
        assert len(r_comp) == kinds.count(1)
        if r_part is not None:
            assert len(r_part) == kinds.count(2)
        # smallest-route removal is deterministic up to ties: the surviving complete tours have the same lengths
        # whenever the picked partial tours have the same lengths
        mine_p = sorted(len(t) for t, k in zip(tours, kinds) if k == 2)
        ref_p = sorted(len(t) for t in r_part) if r_part is not None else None
        if ref_p == mine_p:
            assert sorted(len(t) - 2 for t in r_comp) == sorted(len(t) for t, k in zip(tours, kinds) if k == 1)
