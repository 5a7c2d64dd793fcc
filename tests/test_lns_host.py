"""Host-side pieces of the LNS driver (jampr_plus_b200/lns.py): Solomon reader (io_utils.py:24-128) and the
destructor (lib/lns/destructor.py:64-205) — invariants the repair step relies on."""
import os

import numpy as np

from jampr_plus_b200.lns import Destructor, read_solomon

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def test_read_solomon_c103():
    inst = read_solomon(os.path.join(DATA, "c103.txt"))
    assert inst.graph_size == 101 and inst.max_vehicle_number == 25
    assert inst.org_service_horizon == 1236.0 and inst.type == "1"
    assert abs(float(inst.tw_frac) - 0.5) < 1e-9            # -> checkpoint family all_1_low (config_lns/model/v1.yaml)
    assert np.allclose(inst.coords[0], [0.40, 0.50]) and np.allclose(inst.coords[1], [0.45, 0.68])
    assert np.isclose(inst.demands[1], 10 / 200) and np.isclose(inst.service_time, 90 / 1236)
    assert np.allclose(inst.tw[0], [0.0, 1.0])
    r = read_solomon(os.path.join(DATA, "r101.txt"))
    assert r.graph_size == 101 and abs(float(r.tw_frac) - 1.0) < 1e-9 and r.org_service_horizon == 230.0


def test_destructor_invariants():
    rng = np.random.default_rng(0)
    for mode in ("random", "random_nodes", "random_cut", "const_cut"):
        ds = Destructor(num_partial=3, rm_partial_mode=mode, rm_complete_mode="random", seed=3)
        for _ in range(50):
            perm = rng.permutation(np.arange(1, 41)).tolist()
            cuts = sorted(rng.choice(np.arange(1, 40), size=7, replace=False).tolist())
            sol = [perm[a:b] for a, b in zip([0] + cuts, cuts + [40])]
            out = ds.random_from_route([list(t) for t in sol])
            closed = [t for t in out if len(t) > 1 and t[0] == 0 and t[-1] == 0]
            opened = [t for t in out if not (len(t) > 1 and t[0] == 0 and t[-1] == 0)]
            assert len(opened) <= 3
            seen = [n for t in out for n in t if n != 0]
            assert len(seen) == len(set(seen)), "customer duplicated by the destroy operator"
            for t in closed:
                assert t[1:-1] in sol, "complete tours must be untouched"
            for t in opened:
                assert 0 not in t
                src = [s for s in sol if set(t) <= set(s)]
                assert src, "a destroyed tour must be a sub-sequence of one original tour"
                it = iter(src[0])
                assert all(n in it for n in t), "order of the kept customers must be preserved"
            # at least one complete tour is removed when any exists (frac_rm_complete_routes > 0, destructor.py:188)
            n_cand = sum(len(t) > 1 for t in sol)
            assert len(closed) <= max(n_cand - min(3, n_cand) - 1, 0)


def test_similarity_selection_matches_reference():
    """export_sol(mode="similarity") selection against vectors produced by the reference's seq_match.py and the
    selection loop of env.py:498-545 (tests/golden/make_selection_golden.py)."""
    import json
    from jampr_plus_b200.selection import select_similarity
    cases = json.load(open(os.path.join(os.path.dirname(DATA), "golden", "selection_similarity.json")))
    assert len(cases) == 3
    for c in cases:
        tours, costs = select_similarity(c["samples"], c["costs"], c["best_idx"], c["num_other"])
        assert tours == c["tours"]
        assert costs == c["sel_costs"]
