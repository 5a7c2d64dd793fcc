"""Tensor-native solution I/O of the LNS loop: device import_sol against the host restatement of the
reference's triple loop, the device destruction operator against its CPU twin (bit-exact, shared random numbers), the
all-device LNS loop (RPEnv.destruct) against solution invariants."""
import json
import os
import warnings

import numpy as np
import pytest
import torch

from oracle.destroy_oracle import destroy_plan
from oracle.replay import env_args, load_case

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
STATE = ("plan", "vflags", "row", "nxt", "cur", "cap", "time", "cttd", "total", "visited", "pred", "succ", "row_last",
         "nvis", "allvis", "nbh", "mask", "status")


def _c103(num_samples=20, precision="fp16"):
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.lns import read_solomon
    from jampr_plus_b200.weights import random_state_dict
    inst = read_solomon(os.path.join(DATA, "c103.txt"))
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=num_samples, inference=True,
                tour_graph_update_step=2, device="cuda")
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=precision)
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    return inst, env, pol


def test_device_import_equals_host_import_on_reference_solution():
    """The solution lists of the reference trace g51_repair (construct -> seeded destroy) through both import paths: every
    state array, the first observation and the new step counter agree bit for bit.  (The trace itself — plan, rows,
    visited, totals recorded from the reference's import_sol — is asserted by tests/test_gpu_parity.py through the
    device path.)"""
    from jampr_plus_b200 import RPEnv
    g = load_case("g51_repair")
    sol = json.loads(str(g["solution_json"]))
    snaps = []
    for path in ("device", "host"):
        env = RPEnv(inference=True, device="cuda", **env_args(g["env_kwargs"]))
        env.load_data(g["instances"], dist_mat=torch.from_numpy(g["static_dist"]))
        env.reset()
        env._step = int(g["prev_step"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            obs = env.import_sol(sol) if path == "device" else env._import_sol_host(sol)
        snaps.append(({k: env.state[k].clone() for k in STATE}, env._step, env._graph_fresh))
    for k in STATE:
        assert torch.equal(snaps[0][0][k], snaps[1][0][k]), k
    assert snaps[0][1:] == snaps[1][1:]
    assert np.array_equal(snaps[0][0]["plan"].cpu().numpy(), g["import_plan"])
    assert np.array_equal(snaps[0][0]["total"].cpu().numpy(), g["import_total"])


def test_device_import_expand_and_given_costs():
    """Non-POMO import with fewer samples than num_samples (_expand_sample_dimension, env.py:1204-1219) and
    caller-supplied costs: solutions exported from a sampled episode, mutilated by the host Destructor."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.lns import Destructor
    from jampr_plus_b200.synth import sample_instances
    from jampr_plus_b200.weights import random_state_dict
    inst = sample_instances(2, graph_size=30, seed=3)
    env = RPEnv(inference=True, device="cuda", max_concurrent_vehicles=2, k_nbh_frac=0.4, pomo=False, num_samples=6)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp32")
    pol.load_state_dict(random_state_dict(0))
    pol.set_decode_type("sampling")
    pol.sampling_seed = 5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env.load_data(inst)
        env.reset()
        rollout(pol, env)
        sols, _ = env.export_sol(num_best=1, num_other=1, mode="random")
        ds = Destructor(num_partial=2, seed=3)
        sol = [ds.destruct(s) for s in sols]
        outs = []
        for path in ("device", "host"):
            e = RPEnv(inference=True, device="cuda", max_concurrent_vehicles=2, k_nbh_frac=0.4, pomo=False, num_samples=6)
            e.load_data(inst)
            e.reset()
            f = e.import_sol if path == "device" else e._import_sol_host
            f(sol)
            a = {k: e.state[k].clone() for k in STATE}
            e.load_data(inst)
            e.reset()
            f(sol, cost=[1.0, 2.0, 3.0, 4.0])
            outs.append((a, {k: e.state[k].clone() for k in STATE}))
    for k in STATE:
        assert torch.equal(outs[0][0][k], outs[1][0][k]), k
        assert torch.equal(outs[0][1][k], outs[1][1][k]), k
    assert outs[0][1]["total"].tolist() == [1.0] * 3 + [2.0] * 3 + [3.0] * 3 + [4.0] * 3
    assert torch.equal(outs[0][0]["plan"][0], outs[0][0]["plan"][2]) and not torch.equal(outs[0][0]["plan"][0], outs[0][0]["plan"][3])


@pytest.mark.parametrize("pmode,cmode,fp,fc", [("random", "random", 0.5, 0.15), ("random_nodes", "smallest", 0.0, 0.3),
                                               ("random_cut", "random", 0.0, 0.0), ("const_cut", "smallest", 0.35, 0.2)])
def test_device_destroy_equals_cpu_twin(pmode, cmode, fp, fc):
    from jampr_plus_b200 import _cabi
    from jampr_plus_b200.driver import rollout
    inst, env, pol = _c103()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env.load_data([inst])
        env.reset()
        rollout(pol, env)
        env.destruct(num_best=2, num_other=3, num_partial=3, rm_partial_mode=pmode, rm_complete_mode=cmode,
                     frac_rm_nodes_from_partial=fp, frac_rm_complete_routes=fc, seed=4242, iteration=7)
    tours, kind = (t.cpu().numpy() for t in env._destroy_buf)
    src = env._last_sources.cpu().numpy().reshape(-1, tours.shape[1], tours.shape[2])
    n_sel = src.shape[0]
    for b in range(env.bs):
        t_ref, k_ref = destroy_plan(src[b % n_sel], b, 4242, 7, 3, _cabi.RM_PARTIAL_MODES[pmode],
                                    _cabi.RM_COMPLETE_MODES[cmode], fp, fc)
        got_k = [int(x) for x in kind[b] if x]
        got_t = [[int(x) for x in row[row > 0]] for row, kd in zip(tours[b], kind[b]) if kd]
        assert got_k == k_ref and got_t == t_ref, b
    # the imported state is consistent with the emitted tours: mutilated tours of at least two customers became the
    # active vehicles, complete tours are finished rows, everything else is unvisited
    st = env.state
    for b in range(env.bs):
        rows = [[int(x) for x in r[r > 0]] for r in st["plan"][b].cpu().numpy()]
        exp = [t for t, kd in zip(*destroy_plan(src[b % n_sel], b, 4242, 7, 3, _cabi.RM_PARTIAL_MODES[pmode],
                                                _cabi.RM_COMPLETE_MODES[cmode], fp, fc)) if kd == 1 or len(t) > 1]
        assert [r for r in rows if r] == exp
        vis = set(np.flatnonzero(st["visited"][b].cpu().numpy()).tolist())
        assert vis == {n for t in exp for n in t}


def test_all_device_lns_loop():
    from jampr_plus_b200.lns import LNS
    inst, env, pol = _c103()
    lns = LNS(inst, pol, env, seed=1234)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plan, cost, true_cost = lns.search_device(num_steps=8)
    assert lns.step == 8 and np.isfinite(cost) and np.isfinite(true_cost)
    inc = [h[1] for h in lns.history]
    assert all(b <= a + 1e-6 for a, b in zip(inc, inc[1:])), "incumbent must never get worse"
    rows = [[int(x) for x in r[r > 0]] for r in plan.cpu().numpy()]
    seen = [n for t in rows for n in t]
    assert len(seen) == len(set(seen)) and set(seen) <= set(range(1, inst.graph_size))
    c = np.asarray(inst.coords, dtype=np.float64) * 100
    for t in rows:
        load, tm, prev = 0.0, 0.0, 0
        for n in t:
            tm += np.floor(10 * np.hypot(*(c[prev] - c[n]))) / 10
            assert tm <= inst.tw[n, 1] * inst.org_service_horizon + 1e-3, "time window violated"
            tm = max(tm, inst.tw[n, 0] * inst.org_service_horizon) + inst.service_time * inst.org_service_horizon
            load += inst.demands[n]
            prev = n
        assert load <= 1.0 + 1e-6, "capacity violated"


def test_edge_cases_of_the_device_solution_io():
    """Empty and degenerate inputs: a solution without tours (everything unvisited), a solution whose every tour is a
    singleton, a partial tour that fills the tour buffer, too many tours (the reference's RuntimeError), destruction
    of solutions with fewer candidate tours than num_partial, local search on an empty / single-tour plan."""
    from jampr_plus_b200 import RPEnv
    from jampr_plus_b200.synth import sample_instances
    inst = sample_instances(1, graph_size=40, seed=8)
    env = RPEnv(inference=True, device="cuda", max_concurrent_vehicles=2, k_nbh_frac=0.5, pomo=False, num_samples=4)
    env.load_data(inst)
    env.reset()
    K_max, L = env.state.K_max, env.state.L
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # no tours at all / only singletons -> the state of a fresh reset (two fresh vehicles, nothing visited)
        for sol in ([[[], [[3]], [[5], [7]], []]],):
            for path in ("device", "host"):
                e = RPEnv(inference=True, device="cuda", max_concurrent_vehicles=2, k_nbh_frac=0.5, pomo=False, num_samples=4)
                e.load_data(inst)
                e.reset()
                (e.import_sol if path == "device" else e._import_sol_host)(sol)
                assert int(e.state["visited"].sum()) == 0 and e._step == 0
                assert e.state["row"].tolist() == [[0, 1]] * 4 and float(e.state["total"].abs().sum()) == 0.0
        # more tours than vehicles: the reference's RuntimeError (env.py:660-662)
        too_many = [[[[0, 1 + t, 0] for t in range(K_max + 1)]] * 4]
        with pytest.raises(RuntimeError):
            env.load_data(inst)
            env.reset()
            env.import_sol(too_many)
        # destruction with a single candidate tour and num_partial = 2, complete-route removal on an empty rest
        env.load_data(inst)
        env.reset()
        plan = torch.zeros(4, K_max, L, dtype=torch.int16, device="cuda")
        plan[:, 0, :5] = torch.tensor([4, 9, 2, 7, 11], dtype=torch.int16)
        plan[:, 1, 0] = 13                                   # a singleton tour: always dropped
        env.state["plan"].copy_(plan)
        env.state["total"].copy_(torch.tensor([1.0, 2.0, 3.0, 4.0], device="cuda"))
        env.destruct(num_best=1, num_other=1, num_partial=2, rm_partial_mode="random_cut", frac_rm_complete_routes=0.5,
                     seed=3, iteration=1)
        tours, kind = (t.cpu().numpy() for t in env._destroy_buf)
        for b in range(4):
            assert kind[b].tolist()[:1] == [2] and not kind[b][1:].any()
            t = tours[b, 0][tours[b, 0] > 0].tolist()
            assert 1 <= len(t) <= 3 and t == [4, 9, 2, 7, 11][: len(t)]
        # local search: empty plan and single tour are fixed points or improve, never break
        env.state["plan"].zero_()
        cost, moves = env.local_search(10, update_totals=False)
        assert float(cost.abs().sum()) == 0.0 and int(moves.sum()) == 0
        env.state["plan"][:, 0, :3] = torch.tensor([4, 9, 2], dtype=torch.int16, device="cuda")
        cost, moves = env.local_search(10, update_totals=False)
        after = env.state["plan"][:, :, :].cpu().numpy()
        for b in range(4):
            assert sorted(after[b][after[b] > 0].tolist()) == [2, 4, 9]
        assert torch.isfinite(cost).all()
