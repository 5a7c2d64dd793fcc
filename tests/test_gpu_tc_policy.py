"""GPU parity of the tensor-core policy step (csrc/policy_tc2.cu / policy_tc.cu: tcgen05 GraphConv layers + fused
decoder) against the golden traces recorded from the reference, under teacher forcing.

Bars come from BASELINE.json's north_star, not from what the kernel happens to deliver:
  * environment side (candidate sets, masks, costs, tour buffers): bit-exact, it is the same fp32 code;
  * log-probabilities, fp16 operands (the product precision): within 1e-3 RELATIVE of the reference,
    |dlogp| <= LOGP_RTOL * max(1, |logp_ref|)   (unit floor: a relative bound on a log-probability near 0 would be a
    bound on rounding noise); measured 5.7e-4 at N = 101 (tools/emulate_tc_precision.py reproduces the kernel's
    roundings on the CPU: 2.5e-3 absolute at |logp| ~ 4.4);
  * greedy decisions: identical wherever the reference's top two log-probabilities are not tied within that
    tolerance, i.e. gap > tol(top1) + tol(top2).
bf16 operands (8-bit significand) miss the tolerance by an order of magnitude (3.3e-2 absolute at N = 101) and are not
a supported product precision (Policy refuses them without allow_reduced_accuracy=True); the bf16 rows below only
characterise that mode with bars taken from its measurement and make no parity claim."""
import warnings

import numpy as np
import pytest
import torch

from oracle.replay import env_args, load_case, start_nodes_of

pytestmark = pytest.mark.gpu

LOGP_RTOL = 1e-3                 # north_star: "log-probabilities ... within 1e-3 relative"
BF16_LOGP_ATOL = 8e-2            # characterisation only (measured 3.3e-2 .. 4e-2), no parity claim
BF16_DECISION_GAP = 8e-2
EMB_ATOL = {"fp16": 4e-3, "bf16": 3e-2}      # h_fin / node_emb against the fp32 CUDA path: an internal quantity of O(1),
                                             # bounded by the 16-bit operand rounding of six GraphConv blocks
CASES = ["g21_pomo", "g101_pomo", "g51_sampling", "g51_upd2", "g31_k2"]


def logp_tol(ref):
    return LOGP_RTOL * np.maximum(1.0, np.abs(ref))


def _make(g, precision, tc_variant=0):
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.weights import random_state_dict
    env = RPEnv(inference=True, device="cuda", **env_args(g["env_kwargs"]))
    env.seed(1)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=precision, allow_reduced_accuracy=True,
                 tc_variant=tc_variant)
    pol.load_state_dict(random_state_dict(int(g["policy_seed"])))
    pol.eval()
    pol.return_logp = True
    env.load_data(g["instances"], dist_mat=torch.from_numpy(g["static_dist"]))
    return env, pol


def test_bf16_is_not_a_product_precision():
    from jampr_plus_b200 import Policy, RPEnv
    with pytest.raises(NotImplementedError, match="1e-3"):
        Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="bf16")


def _sn(g):
    sn = start_nodes_of(g)
    return None if sn is None else sn.to(torch.int32)


@pytest.mark.parametrize("precision,variant", [("fp16", 2), ("fp16", 3), ("bf16", 3)])
@pytest.mark.parametrize("name", CASES)
def test_teacher_forced_tc(name, precision, variant):
    """variant 2 = decoder fused per rollout (what small batches run), 3 = tcgen05 decoder over 128 rollouts (large
    batches): both against the reference trace, same bars."""
    g = load_case(name)
    env, pol = _make(g, precision, variant)
    envr, polr = _make(g, "fp32")
    obs, obr = env.reset(start_nodes=_sn(g)), envr.reset(start_nodes=_sn(g))
    T = g["trace_action"].shape[0]
    worst, worst_rel, flips, n_dec, n_decisive = 0.0, 0.0, 0, 0, 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for t in range(T):
            assert np.array_equal(obs.nbh.cpu().numpy(), g["trace_nbh"][t]), f"nbh step {t}"
            assert np.array_equal(obs.nbh_mask.cpu().numpy(), g["trace_mask"][t]), f"mask step {t}"
            pol(obs)
            polr(obr)
            lp = pol.log_probs(env).cpu().numpy()
            ref = g["trace_logp"][t]
            fin = np.isfinite(ref)
            assert np.array_equal(np.isfinite(lp), fin), f"support step {t}"
            err = np.abs(lp - ref)[fin]
            worst = max(worst, float(err.max()))
            worst_rel = max(worst_rel, float((err / np.maximum(1.0, np.abs(ref[fin]))).max()))
            if precision == "fp16":
                assert bool((err <= logp_tol(ref[fin])).all()), (t, float(err.max()), worst_rel)
            else:
                assert float(err.max()) <= BF16_LOGP_ATOL, (t, float(err.max()))
            if env._graph_fresh:
                assert float((env.state["h_fin"] - envr.state["h_fin"]).abs().max()) <= EMB_ATOL[precision], t
                assert float((env.state["ne"] - envr.state["ne"]).abs().max()) <= EMB_ATOL[precision], t
            if g["decode"] == "greedy":
                srt = np.sort(ref, -1)
                gap = srt[:, -1] - srt[:, -2]
                decisive = gap > (logp_tol(srt[:, -1]) + logp_tol(srt[:, -2]) if precision == "fp16" else BF16_DECISION_GAP)
                mine = env.state["action"].cpu().numpy()
                same = (mine == g["trace_action"][t]).all(-1)
                assert same[decisive].all(), f"decisive greedy action differs at step {t}"
                flips += int((~same).sum())
                n_dec += same.size
                n_decisive += int(decisive.sum())
            act = torch.from_numpy(g["trace_action"][t].astype(np.int64))
            obs, cost, done, _ = env.step(act)
            obr, _, _, _ = envr.step(act)
            if not done:
                assert np.array_equal(cost.cpu().numpy(), g["trace_cost"][t]), f"cost step {t}"
    assert done
    assert np.array_equal(env.tour_plan.cpu().numpy(), g["final_tour_plan"])
    if g["decode"] == "greedy":
        assert n_decisive > (0.5 * n_dec if precision == "fp16" else 0), "the decisive set must not be (nearly) empty"
    print(f"{name}/{precision}: worst |dlogp| {worst:.2e} (relative {worst_rel:.2e}), {flips}/{n_dec} near-tie decisions "
          f"differ, {n_decisive} decisive")


@pytest.mark.parametrize("precision,variant", [("fp16", 2), ("fp16", 3)])
@pytest.mark.parametrize("name", ["g21_pomo", "g31_k2", "g51_upd2"])
def test_fused_rollout_tc(name, precision, variant):
    """Free-running device loop on the tensor-core path: every rollout whose reference decisions are all decisive
    reproduces the reference tours; the step API and the fused loop agree bit for bit with each other."""
    from jampr_plus_b200.driver import rollout
    g = load_case(name)
    env, pol = _make(g, precision, variant)
    pol.return_logp = False
    srt = np.sort(g["trace_logp"], -1)
    decided = ((srt[..., -1] - srt[..., -2]) > logp_tol(srt[..., -1]) + logp_tol(srt[..., -2])).all(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        env.reset(start_nodes=_sn(g))
        total, info = rollout(pol, env)
        plan_fused = env.tour_plan.clone()
        assert info["done_step"] >= 0
        same = (plan_fused.cpu().numpy() == g["final_tour_plan"]).all(axis=(1, 2))
        assert same[decided].all()           # often vacuous: ~N decisions per rollout, one near-tie is enough
        assert same.mean() >= 0.6, same      # measured: >= 21/24 (g21), teacher-forced flips are < 0.5 % of decisions
        env.load_data(g["instances"], dist_mat=torch.from_numpy(g["static_dist"]))
        obs = env.reset(start_nodes=_sn(g))
        done = False
        while not done:
            action, _, _ = pol(obs)
            obs, _, done, _ = env.step(action)
    assert torch.equal(env.tour_plan, plan_fused)
    assert torch.equal(env.total_cost, total)


@pytest.mark.parametrize("variant", [2, 3])
def test_tc_sampling_uses_given_uniforms(variant):
    """Sampling decode on the tensor-core path: the action is the inverse-CDF pick of the kernel's own
    log-probabilities for the caller's uniforms (same rule as the fp32 path / oracle.select_inverse_cdf)."""
    from oracle.policy_oracle import select_inverse_cdf
    g = load_case("g51_sampling")
    env, pol = _make(g, "fp16", variant)
    pol.set_decode_type("sampling")
    obs = env.reset()
    gen = torch.Generator().manual_seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for t in range(20):
            u = torch.rand(env.bs, generator=gen)
            action, ll, ent = pol(obs, uniforms=u)
            logp = pol.log_probs(env).cpu()
            act_o, ll_o = select_inverse_cdf(logp, obs.nbh.cpu().long(), u)
            assert torch.equal(action.cpu(), act_o), f"step {t}"
            assert torch.equal(ll.cpu(), ll_o)
            assert torch.isfinite(ent)
            obs, _, done, _ = env.step(action)
            if done:
                break


@pytest.mark.parametrize("n_customers,k_frac,K,P,variant",
                         [(119, 0.25, 3, 4, 0), (127, 0.1, 4, 2, 0), (60, 0.4, 1, 4, 0), (12, 0.5, 2, 4, 0),
                          # the 128-rollout decoder (variant 3) on its corner shapes: one vehicle, four vehicles with
                          # 180 actions (twelve 16-row chunks per rollout), a tiny graph, more rollouts than one tile
                          (60, 0.4, 1, 4, 3), (100, 0.45, 4, 2, 3), (12, 0.5, 2, 4, 3), (40, 0.3, 3, 50, 3)])
def test_tc_matches_fp32_path_other_shapes(n_customers, k_frac, K, P, variant):
    """Shapes outside the golden fixtures (graph sizes up to the 128-row tile, 1..4 concurrent vehicles; N > 104 runs
    the one-CTA-per-SM tiling, the others the two-CTA tiling): the tensor-core step against the fp32 CUDA path,
    teacher-forced with the fp32 path's greedy actions."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.synth import sample_instances
    from jampr_plus_b200.weights import random_state_dict
    inst = sample_instances(3, graph_size=n_customers, seed=5)
    envs, pols = [], []
    for prec in ("fp32", "fp16"):
        env = RPEnv(max_concurrent_vehicles=K, k_nbh_frac=k_frac, pomo=True, pomo_nbh_frac=0.6, num_samples=P,
                    inference=True, device="cuda")
        env.seed(3)
        pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=prec, tc_variant=variant)
        pol.load_state_dict(random_state_dict(1))
        pol.return_logp = True
        env.load_data(inst)
        envs.append(env)
        pols.append(pol)
    obs0 = envs[0].reset()
    obs1 = envs[1].reset(start_nodes=envs[0]._start_nodes)
    worst, done, steps = 0.0, False, 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        while not done:
            assert torch.equal(obs0.nbh, obs1.nbh) and torch.equal(obs0.nbh_mask, obs1.nbh_mask)
            action, _, _ = pols[0](obs0)
            pols[1](obs1)
            l0, l1 = pols[0].log_probs(envs[0]), pols[1].log_probs(envs[1])
            fin = torch.isfinite(l0)
            assert torch.equal(fin, torch.isfinite(l1))
            worst = max(worst, float((l0[fin] - l1[fin]).abs().max()))
            obs0, c0, done, _ = envs[0].step(action)
            obs1, c1, _, _ = envs[1].step(action)
            assert torch.equal(c0, c1)
            steps += 1
    if variant == 3:
        assert worst <= 6e-3, worst  # split-operand decoder: the 16-bit GNN operands are what is left (2.5e-3 at N = 101)
    assert worst <= 2e-2, worst      # against the fp32 CUDA path on shapes without a reference trace: a smoke bound on
    assert steps > n_customers // 2  # the other tilings (the traced fixtures above carry the parity claim)


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_tc_rollout_is_deterministic_under_load(precision):
    """8192 rollouts (many waves of two co-resident CTAs per SM, overlaid shared-memory regions, TMEM reuse): two
    free-running episodes from the same start must agree bit for bit — tours, totals and per-step log-likelihood
    sums.  A race between the kernel's phases shows up here as a run-to-run difference."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.synth import sample_instances
    from jampr_plus_b200.weights import random_state_dict
    inst = sample_instances(512, graph_size=100, seed=9)
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=16, inference=True, device="cuda")
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=precision, allow_reduced_accuracy=True)
    pol.load_state_dict(random_state_dict(2))
    pol.eval()
    outs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for _ in range(2):
            env.load_data(inst)
            env.seed(4)
            env.reset()
            total, info = rollout(pol, env)
            outs.append((env.tour_plan.clone(), total.clone(), info["done_step"]))
    assert outs[0][2] == outs[1][2] and outs[0][2] >= 0
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert torch.isfinite(outs[0][1]).all()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_seeded_sampling_replays_on_cpu(precision, variant=3):
    """Sampling without caller-supplied uniforms: the kernel draws from jampr_uniform(seed, rollout, step)
    (csrc/common.cuh); oracle/rng.py is its CPU twin, so the episode can be replayed: the inverse-CDF pick of the
    kernel's own log-probabilities at the twin's uniforms must equal the kernel's action, every step."""
    from oracle.policy_oracle import select_inverse_cdf
    from oracle.rng import uniforms
    g = load_case("g51_sampling")
    env, pol = _make(g, precision, variant)
    pol.set_decode_type("sampling")
    pol.sampling_seed = 987654321
    obs = env.reset()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for t in range(25):
            step_no = env._step
            action, ll, _ = pol(obs)
            logp = pol.log_probs(env).cpu()
            u = torch.from_numpy(uniforms(pol.sampling_seed, env.bs, step_no))
            act_o, ll_o = select_inverse_cdf(logp, obs.nbh.cpu().long(), u)
            assert torch.equal(action.cpu(), act_o), f"step {t}"
            assert torch.equal(ll.cpu(), ll_o)
            obs, _, done, _ = env.step(action)
            if done:
                break


def test_two_policies_sharing_one_env_reencode():
    """The node-encoder output lives in the env's buffers and is precision specific: alternating an fp32 and an fp16
    policy on one env must give each its own static tables (stale encode_static key)."""
    g = load_case("g21_pomo")
    env, p16 = _make(g, "fp16")
    _, p32 = _make(g, "fp32")
    env2, q32 = _make(g, "fp32")
    obs = env.reset(start_nodes=_sn(g))
    obs2 = env2.reset(start_nodes=_sn(g))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        p32(obs)
        ref = p32.log_probs(env).clone()
        p16(obs)
        p32(obs)                      # must re-encode: the tables were last written by the fp16 policy
        again = p32.log_probs(env).clone()
        q32(obs2)
    assert torch.equal(ref, again)
    assert torch.equal(ref, q32.log_probs(env2))
