"""Pin the CPU oracle against traces recorded from the unmodified reference (tests/golden/*.npz).

Teacher forcing: the oracle environment is driven with the actions the reference took; every step
its candidate sets, feasibility masks, tour features and step costs must equal the reference's
bit for bit, and the oracle policy's log-probabilities must agree to LOGP_ATOL."""
import numpy as np
import pytest
import torch

from oracle.replay import CASES, load_case, make_oracle, start_nodes_of
from oracle.policy_oracle import select_greedy

LOGP_ATOL = 2e-5     # fp32 op-order noise between two CPU restatements of the same network


@pytest.mark.parametrize("name", CASES)
def test_static_tables(name):
    g = load_case(name)
    env, _ = make_oracle(g)
    assert env.K_max == int(g["static_k_max"])
    # the formula (correctly rounded sqrt) against the reference's CPU matrix (MKL sqrt): equal except for
    # isolated pairs that sit exactly one 0.1-step lower in the reference (see env_oracle.dimacs_dist)
    ref = g["static_dist"]
    mine = env.dist_formula.numpy()
    off = mine != ref
    assert off.mean() <= 1e-3
    hz = g["inst_org_service_horizon"].astype(np.float32)[:, None, None]
    assert np.all(np.abs((mine - ref) * hz)[off] <= 0.1001)
    assert np.array_equal(env.tw.numpy(), g["static_tw_fixed"])
    assert np.array_equal(env.knn[:, 1:, :].numpy().astype(np.int16), g["static_knn"])
    if "static_ordered_idx" in g:
        assert np.array_equal(env.order.numpy().astype(np.int16), g["static_ordered_idx"])


@pytest.mark.parametrize("name", CASES)
def test_teacher_forced_episode(name):
    g = load_case(name)
    env, pol = make_oracle(g)
    env.reset(start_nodes=start_nodes_of(g))
    assert np.array_equal(env.plan.numpy(), g["start_plan"])
    T = g["trace_action"].shape[0]
    worst = 0.0
    for t in range(T):
        nbh, mask = env.observe()
        assert np.array_equal(nbh.numpy().astype(np.int16), g["trace_nbh"][t]), f"nbh step {t}"
        assert np.array_equal(mask.numpy(), g["trace_mask"][t]), f"mask step {t}"
        assert np.array_equal(env.tour_features().numpy(), g["trace_tour_features"][t]), f"tour features step {t}"
        assert env.graph_fresh == (int(g["trace_n_tour_edges"][t]) >= 0), f"graph refresh rule step {t}"
        logp = pol.forward(env, nbh, mask).numpy()
        ref = g["trace_logp"][t]
        assert np.array_equal(np.isfinite(logp), np.isfinite(ref))
        fin = np.isfinite(ref)
        worst = max(worst, float(np.abs(logp[fin] - ref[fin]).max()))
        cost, done = env.step(torch.from_numpy(g["trace_action"][t].astype(np.int64)))
        assert np.array_equal(cost.numpy(), g["trace_cost"][t]), f"cost step {t}"
        assert np.array_equal(env.total.numpy(), g["trace_total"][t]), f"total step {t}"
    assert done
    assert worst <= LOGP_ATOL, worst
    assert np.array_equal(env.plan.numpy(), g["final_tour_plan"])
    assert np.array_equal(env.total.numpy(), g["final_total_cost"])
    assert np.array_equal(env.k_used().numpy(), g["final_k_used"])
    assert np.array_equal(env.visited.numpy(), g["final_visited"])
    p, s = env.links_from_plan()
    assert torch.equal(p, env.pred) and torch.equal(s, env.succ)


@pytest.mark.parametrize("name", ["g21_pomo", "g31_k2", "g51_upd2"])
def test_free_running_greedy_matches_reference_tours(name):
    """Without teacher forcing the oracle must take the reference's decisions wherever the reference's
    top-2 log-prob gap exceeds the tolerance; on these fixtures that means identical tours."""
    g = load_case(name)
    env, pol = make_oracle(g)
    env.reset(start_nodes=start_nodes_of(g))
    done, t = False, 0
    while not done:
        nbh, mask = env.observe()
        act, _ = select_greedy(pol.forward(env, nbh, mask), nbh)
        ref_lp = np.sort(g["trace_logp"][t], -1)
        decided = (ref_lp[:, -1] - ref_lp[:, -2]) > 10 * LOGP_ATOL
        assert np.array_equal(act.numpy()[decided], g["trace_action"][t].astype(np.int64)[decided]), f"step {t}"
        _, done = env.step(torch.from_numpy(g["trace_action"][t].astype(np.int64)))
        t += 1
    assert np.array_equal(env.plan.numpy(), g["final_tour_plan"])


def test_node_encoder_matches_reference():
    g = load_case("g101_pomo")
    env, pol = make_oracle(g)
    emb = pol.encode_nodes(env)[0].numpy()
    assert np.abs(emb - g["static_node_emb0"]).max() <= 2e-5


def test_repair_step_import_sol():
    """LNS repair step: the oracle's import_sol (env.py:575-706) and the repair rollout that follows, against the
    reference trace (construct -> seeded destroy -> import_sol -> greedy repair, tests/golden/make_golden.py)."""
    import json
    g = load_case("g51_repair")
    env, pol = make_oracle(g)
    env.reset(start_nodes=None if not env.pomo else torch.zeros(env.bs, env.K, dtype=torch.long) + 1)
    env.step_no = int(g["prev_step"])                 # the construction episode that preceded the import
    env.import_sol(json.loads(str(g["solution_json"])))
    assert np.array_equal(env.plan.numpy(), g["import_plan"])
    assert np.array_equal(env.row.numpy(), g["import_rows"])
    assert np.array_equal(env.visited.numpy(), g["import_visited"])
    assert np.array_equal(env.total.numpy(), g["import_total"])
    assert env.step_no == int(g["import_step"])
    T = g["trace_action"].shape[0]
    worst = 0.0
    for t in range(T):
        nbh, mask = env.observe()
        assert np.array_equal(nbh.numpy().astype(np.int16), g["trace_nbh"][t]), f"nbh step {t}"
        assert np.array_equal(mask.numpy(), g["trace_mask"][t]), f"mask step {t}"
        assert np.array_equal(env.tour_features().numpy(), g["trace_tour_features"][t]), f"tour features step {t}"
        assert env.graph_fresh == (int(g["trace_n_tour_edges"][t]) >= 0), f"graph refresh rule step {t}"
        logp = pol.forward(env, nbh, mask).numpy()
        ref = g["trace_logp"][t]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(logp), fin)
        worst = max(worst, float(np.abs(logp[fin] - ref[fin]).max()))
        cost, done = env.step(torch.from_numpy(g["trace_action"][t].astype(np.int64)))
        if not done:
            assert np.array_equal(cost.numpy(), g["trace_cost"][t]), f"cost step {t}"
            assert np.array_equal(env.total.numpy(), g["trace_total"][t]), f"total step {t}"
        else:
            np.testing.assert_allclose(env.total.numpy(), g["trace_total"][t], rtol=1e-6)
    assert done and worst <= LOGP_ATOL, worst
    assert np.array_equal(env.plan.numpy(), g["final_tour_plan"])
    assert env.step_no == int(g["final_step"])
