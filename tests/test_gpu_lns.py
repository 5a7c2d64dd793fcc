"""LNS driver on the GPU path (BASELINE.json configs 1 / 5 without local search): construct -> destroy -> import_sol
-> repair on Solomon C103, checked by solution properties (the shipped checkpoint and OR-tools are absent, so cost
levels are not comparable with published ones)."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _setup(precision="fp16", num_samples=20):
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.lns import LNS, read_solomon
    from jampr_plus_b200.weights import random_state_dict
    inst = read_solomon(os.path.join(DATA, "c103.txt"))
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=num_samples, inference=True,
                tour_graph_update_step=2, device="cuda")                    # config_lns/env_cfg_overrides/K1.yaml
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=precision)
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    return inst, env, pol, LNS(inst, pol, env, seed=1234)


def _check(sol, inst):
    seen = [n for t in sol for n in t]
    assert len(seen) == len(set(seen)) and 0 not in seen
    assert set(seen) <= set(range(1, inst.graph_size))


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_lns_loop_c103(precision):
    inst, env, pol, lns = _setup(precision)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        best, cost, true_cost = lns.search(num_steps=6)
    _check(best, inst)
    inc = [h[1] for h in lns.history]
    assert all(b <= a + 1e-6 for a, b in zip(inc, inc[1:])), "incumbent must never get worse"
    assert lns.step == 6 and np.isfinite(cost) and np.isfinite(true_cost)
    # capacity / time windows of the incumbent, recomputed on the host in original units
    c = np.asarray(inst.coords, dtype=np.float64) * 100
    for t in best:
        load, tm, prev = 0.0, 0.0, 0
        for n in t:
            d = np.floor(10 * np.hypot(*(c[prev] - c[n]))) / 10
            tm += d
            assert tm <= inst.tw[n, 1] * inst.org_service_horizon + 1e-3, "time window violated"
            tm = max(tm, inst.tw[n, 0] * inst.org_service_horizon) + inst.service_time * inst.org_service_horizon
            load += inst.demands[n]
            prev = n
        assert load <= 1.0 + 1e-6, "capacity violated"


def test_graph_replay_equals_eager():
    """rollout(graph=True): the captured CUDA graph of the whole episode reproduces the eager run bit for bit."""
    from jampr_plus_b200.driver import rollout
    inst, env, pol, _ = _setup("fp16", num_samples=20)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env.load_data([inst])
        env.reset()
        sn = env._start_nodes.clone()
        total_e, _ = rollout(pol, env)
        plan_e, total_e = env.tour_plan.clone(), total_e.clone()
        for _ in range(3):                      # first call captures, later calls replay
            env.load_data([inst])
            env.reset(start_nodes=sn)
            total_g, info = rollout(pol, env, graph=True)
            assert info["done_step"] >= 0
            assert torch.equal(env.tour_plan, plan_e) and torch.equal(total_g, total_e)
