"""Device local search (csrc/local_search.cu, SURVEY.md §8 f4: the stand-in for the reference's OR-tools step) against
its CPU twin (oracle/ls_oracle.py: identical tours, costs and move counts) and against solution invariants at the LNS
sizes; the LNS loop with the search switched on."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle.ls_oracle import local_search as ls_twin
from oracle.ls_oracle import sim
from oracle.replay import env_args, load_case, start_nodes_of

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _finished_env(case):
    from jampr_plus_b200 import RPEnv
    g = load_case(case)
    env = RPEnv(inference=True, device="cuda", **env_args(g["env_kwargs"]))
    env.load_data(g["instances"], dist_mat=torch.from_numpy(g["static_dist"]))
    sn = start_nodes_of(g)
    env.reset(start_nodes=None if sn is None else sn.to(torch.int32))
    env.state["plan"].copy_(torch.from_numpy(g["final_tour_plan"]).cuda())
    env.state["visited"].copy_(torch.from_numpy(g["final_visited"].astype(np.uint8)).cuda())
    return g, env


@pytest.mark.parametrize("case,iters", [("g21_pomo", 50), ("g31_k2", 6)])
def test_device_search_equals_cpu_twin(case, iters):
    g, env = _finished_env(case)
    before = env.state["plan"].cpu().numpy().copy()
    tc_before = env.true_cost().cpu().numpy()
    cost, moves = env.local_search(max_iters=iters)
    after = env.state["plan"].cpu().numpy()
    P = env.num_samples
    dist, tw, dem, srv = (env._dist_mat.cpu().numpy(), env.tw.cpu().numpy(), env.demands.cpu().numpy(),
                          env.service_time.cpu().numpy())
    improved = 0
    for b in range(env.bs):
        i = b // P
        ref_plan, ref_cost, ref_moves = ls_twin(before[b], dist[i], tw[i], dem[i], srv[i], iters)
        assert np.array_equal(after[b], ref_plan), b
        assert int(moves[b]) == ref_moves and float(cost[b]) == float(ref_cost), b
        improved += ref_moves > 0
        # same customers as before, each once
        assert sorted(before[b][before[b] > 0].tolist()) == sorted(after[b][after[b] > 0].tolist())
    assert improved > 0
    tc_after = env.true_cost().cpu().numpy()
    assert (tc_after <= tc_before + 1e-6).all()
    assert np.array_equal(env.total_cost.cpu().numpy(), tc_after)


def test_search_result_is_feasible_for_the_environment():
    """C103 (N = 101): the improved tours, imported as the reference's solution lists and re-simulated by the
    environment's import (env.py:575-706), are complete tours whose recomputed cost matches; every tour passes the
    environment's capacity / window tests; the search never increases a solution's cost."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.lns import read_solomon
    from jampr_plus_b200.weights import random_state_dict
    inst = read_solomon(os.path.join(DATA, "c103.txt"))
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=20, inference=True,
                tour_graph_update_step=2, device="cuda")
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
    pol.load_state_dict(random_state_dict(0))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env.load_data([inst])
        env.reset()
        rollout(pol, env)
        tc0 = env.true_cost().clone()
        cost, moves = env.local_search(max_iters=400)
        tc1 = env.true_cost()
    assert int(moves.min()) > 0 and bool((tc1 <= tc0 + 1e-6).all())
    assert float((tc0 - tc1).mean()) > 0.05 * float(tc0.mean()), "random-weight tours leave a lot to improve"
    dist, tw, dem, srv = (env._dist_mat[0].cpu().numpy(), env.tw[0].cpu().numpy(), env.demands[0].cpu().numpy(),
                          float(env.service_time[0]))
    plan = env.tour_plan.cpu().numpy()
    vis = env._visited.cpu().numpy()
    ttd = env.time_to_depot[0].cpu().numpy()
    for b in range(env.bs):
        tot = np.float32(0)
        for r in plan[b]:
            t = [int(x) for x in r[r > 0]]
            c = sim(dist, tw, dem, srv, t)
            assert np.isfinite(c), "a tour fails the environment's capacity / window tests"
            tot = np.float32(tot + c)
        assert abs(float(tot) - float(cost[b])) <= 1e-5 * max(1.0, float(tot))
        unserved = [n for n in range(1, env.graph_size) if not vis[b, n]]
        np.testing.assert_allclose(float(tc1[b]), float(tot) + 2 * float(ttd[unserved].sum()), rtol=1e-5)


def test_lns_with_local_search_beats_lns_without():
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.lns import LNS, read_solomon
    from jampr_plus_b200.weights import random_state_dict
    inst = read_solomon(os.path.join(DATA, "c103.txt"))
    res = {}
    for ls in (0, 1):
        env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=20, inference=True,
                    tour_graph_update_step=2, device="cuda")
        pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
        pol.load_state_dict(random_state_dict(0))
        lns = LNS(inst, pol, env, seed=1234, selection_mode="random", ls_step=ls)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            plan, cost, true_cost = lns.search_device(num_steps=10)
        inc = [h[2] for h in lns.history]
        assert all(b <= a + 1e-3 for a, b in zip(inc, inc[1:]))
        res[ls] = true_cost
    assert res[1] < res[0], res
