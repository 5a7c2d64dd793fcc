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
    """The benchmarked configuration itself (BASELINE config 2: the reference generator's all_1_low instances x 16 POMO
    samples; no log-probability output, so the kernel runs without the embedding cache stores and with the
    idle-rollout skip) on the fp16 tensor-core path against the fp32 CUDA-core path (the 1e-5 anchor):
      1. teacher forced on 512 instances (8192 rollouts, ~1 M decisions): both environments follow the fp32 policy's
         greedy actions; at every step the fp16 policy, in the production configuration, must pick the same action
         wherever the fp32 top-2 log-prob gap exceeds the two 1e-3-relative tolerances (the decisive set is most of the
         decisions, asserted), and its log-likelihood of that action must be within the tolerance;
      2. free running, all 4096 instances (65,536 rollouts) through the fused device loop replayed as a CUDA graph:
         the replay reproduces the eager episode bit for bit, and the cost level (mean per-instance best cost) agrees
         with the fp32 path's free-running episode on the 512-instance slice within 1e-3 relative — a flipped near-tie
         changes a tour, not the cost level."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.weights import random_state_dict
    host, kv = _load_workload("workload_all_1_low_4096.npz")
    I, P, Iref = 4096, 16, 512
    kw = dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=P, inference=True, device="cuda")

    def make(prec, logp):
        e = RPEnv(**kw)
        e.seed(1235)
        p = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=prec)
        p.load_state_dict(random_state_dict(0))
        p.eval()
        p.return_logp = logp
        return e, p

    # ---- 1. teacher forced, production kernel configuration on the fp16 side
    env16, pol16 = make("fp16", False)
    env32, pol32 = make("fp32", True)
    sl = {k: v[:Iref].cuda() for k, v in host.items()}
    env16.load_arrays(sl, kv)
    env32.load_arrays(sl, kv)
    n_dec = n_live = n_flip = 0
    worst_ll = 0.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        o32 = env32.reset()
        sn_slice = env32._start_nodes.clone()
        o16 = env16.reset(start_nodes=sn_slice)
        done = False
        while not done:
            a32, ll32, _ = pol32(o32)
            a16, ll16, _ = pol16(o16)
            top = pol32.log_probs(env32).topk(2, dim=-1).values
            tol = 1e-3 * (top[:, 0].abs().clamp_min(1.0) + top[:, 1].abs().clamp_min(1.0))
            live = ~(env32.state["allvis"].bool() & (env32.state["cur"] == 0).all(-1))   # idle rollouts decide nothing
            decisive = live & (((top[:, 0] - top[:, 1]) > tol) | ~torch.isfinite(top[:, 1]))
            same = (a16 == a32).all(-1)
            assert bool(same[decisive].all()), "a decisive greedy action differs between the fp16 and the fp32 path"
            err = (ll16 - ll32).abs().view(-1)[decisive & same]
            bound = 1e-3 * ll32.abs().view(-1).clamp_min(1.0)[decisive & same]
            assert bool((err <= bound).all()), float((err / bound).max())
            worst_ll = max(worst_ll, float((err / bound).max()) if err.numel() else 0.0)
            n_dec += int(decisive.sum())
            n_live += int(live.sum())
            n_flip += int((live & ~same).sum())
            o32, c32, done, _ = env32.step(a32)
            o16, c16, _, _ = env16.step(a32)
            assert torch.equal(c32, c16)
    assert n_dec > 0.8 * n_live, (n_dec, n_live)
    assert torch.equal(env16.tour_plan, env32.tour_plan)
    # ---- 2. free running at full size, CUDA-graph replay
    env, pol = make("fp16", False)
    dev_arrays = {k: v.cuda() for k, v in host.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        env.load_arrays(dev_arrays, kv)
        env.reset()
        sn = env._start_nodes.clone()
        total, info = rollout(pol, env, graph=True)
        assert info["done_step"] >= 0
        total = total.clone()
        plan = env.tour_plan.clone()
        env.load_arrays(dev_arrays, kv)
        env.reset(start_nodes=sn)
        total_b, _ = rollout(pol, env, graph=True)          # second call replays the captured graph
        assert torch.equal(total, total_b) and torch.equal(plan, env.tour_plan)
        best16 = env.gather_best()[0][:Iref].clone()
        assert torch.isfinite(env.true_cost()).all()
        envr, polr = make("fp32", False)
        envr.load_arrays(sl, kv)
        envr.reset(start_nodes=sn[: Iref * P])
        rollout(polr, envr)
        best32 = envr.gather_best()[0]
    mean_rel = float((best16.mean() - best32.mean()).abs() / best32.mean())
    frac_same = float((best16 == best32).float().mean())
    print(f"fp16 vs fp32, production configuration: {n_dec} decisive of {n_live} live decisions all identical "
          f"({n_flip} near-tie flips), log-likelihood error up to {worst_ll:.2f} of the 1e-3 tolerance; free running: "
          f"per-instance best cost identical for {frac_same:.3f} of the instances, mean differs by {mean_rel:.2e} relative")
    assert mean_rel <= 1e-3, mean_rel
