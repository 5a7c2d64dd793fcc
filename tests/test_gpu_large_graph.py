"""Large-graph path (N > 128: several 128-row tiles per graph, csrc/gnn_large.cu) against the CPU oracle: node
encoder output, per-step node embeddings and log-probabilities under teacher forcing, environment side bit-exact.
BASELINE config 4 (Gehring-Homberger, 1000 customers) is this path at N = 1001; the oracle comparison runs at
N = 301 (three tiles, the last one partial) where the CPU restatement finishes in seconds, plus a smoke run at 1001."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

EMB_ATOL = 6e-3      # fp16 operands, three (node encoder) + six (tour encoder) GraphConv blocks, values O(1)
LOGP_ATOL = 5e-2


def _pair(n_customers, k_frac, n_inst=2, P=2, seed=3):
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.synth import sample_instances
    from jampr_plus_b200.weights import random_state_dict
    from oracle.env_oracle import OracleEnv
    from oracle.policy_oracle import OraclePolicy
    inst = sample_instances(n_inst, graph_size=n_customers, seed=seed)
    kw = dict(max_concurrent_vehicles=3, k_nbh_frac=k_frac, pomo=True, pomo_single_start_node=True, num_samples=P)
    env = RPEnv(inference=True, device="cuda", **kw)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
    pol.load_state_dict(random_state_dict(0))
    pol.return_logp = True
    oenv = OracleEnv(**kw)
    oenv.load(inst)
    env.load_data(inst, dist_mat=oenv.dist)
    opol = OraclePolicy(random_state_dict(0))
    return inst, env, pol, oenv, opol


def test_large_graph_teacher_forced_vs_oracle():
    from oracle.policy_oracle import select_greedy
    inst, env, pol, oenv, opol = _pair(300, 0.1)
    assert env.graph_size == 301 and env.k_nbh_size == 31
    obs = env.reset()
    oenv.reset()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for t in range(10):
            nbh, mask = oenv.observe()
            assert torch.equal(obs.nbh.cpu().long(), nbh), f"nbh step {t}"
            assert torch.equal(obs.nbh_mask.cpu(), mask), f"mask step {t}"
            pol(obs)
            tour_emb, node_emb = opol.tour_encode(oenv)
            lo = opol.decode(tour_emb, node_emb, nbh, mask)
            if t == 0:
                se = env._static_emb.cpu()
                assert (se - opol.static_emb).abs().max() <= EMB_ATOL, "node encoder (static_emb)"
            ne = env.state["ne"].cpu()
            assert (ne - node_emb).abs().max() <= EMB_ATOL, f"node_emb step {t}"
            lp = pol.log_probs(env).cpu()
            fin = torch.isfinite(lo)
            assert torch.equal(fin, torch.isfinite(lp))
            assert (lp[fin] - lo[fin]).abs().max() <= LOGP_ATOL, f"log-probs step {t}"
            act, _ = select_greedy(lo, nbh)
            obs, cost, done, _ = env.step(act)
            co, _ = oenv.step(act)
            assert torch.equal(cost.cpu(), co), f"cost step {t}"
    assert torch.equal(env.tour_plan.cpu(), oenv.plan)


def test_gh1000_shape_rollout_smoke():
    """N = 1001, k = 51 (K6 overrides: k_nbh_frac 0.05, tour_graph_update_step 4), 2 instances x 2 samples: the fused
    device loop runs a prefix of the episode; solutions so far respect capacity / windows; no status errors."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.synth import sample_instances
    from jampr_plus_b200.weights import random_state_dict
    inst = sample_instances(2, graph_size=1000, seed=9)
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.05, pomo=True, num_samples=2, inference=True,
                tour_graph_update_step=4, device="cuda")
    env.seed(1)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
    pol.load_state_dict(random_state_dict(0))
    env.load_data(inst)
    env.reset()
    assert env.k_nbh_size == 51 and env.graph_size == 1001
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        total, info = rollout(pol, env, max_steps=120)
    assert info["steps_enqueued"] == 120 and info["done_step"] < 0
    plan = env.tour_plan.cpu().numpy()
    vis = env._visited.cpu().numpy()
    assert vis[:, 1:].sum(-1).min() >= 60, "every rollout must have served customers by now"
    for b in range(env.bs):
        seen = plan[b][plan[b] > 0]
        assert len(seen) == len(set(seen.tolist()))
    assert torch.isfinite(total).all()


def test_gehring_homberger_files():
    """Real 1000-customer files through the reader and the device path: C1_10_1 loads and rolls out (K6 overrides);
    R1_10_1 raises the reference's own time-window fix-up assertion (env.py:408), like 2 of the 60 N1000 files do in
    the reference."""
    import os
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.lns import read_solomon
    from jampr_plus_b200.weights import random_state_dict
    data = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
    c1 = read_solomon(os.path.join(data, "C1_10_1.TXT"))
    assert c1.graph_size == 1001 and c1.max_vehicle_number == 250
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.05, pomo=True, num_samples=4, inference=True,
                tour_graph_update_step=4, device="cuda")
    env.seed(1)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
    pol.load_state_dict(random_state_dict(0))
    env.load_data([c1])
    assert env.max_vehicle_number == 255 and env.k_nbh_size == 51
    env.reset()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        total, info = rollout(pol, env)
    assert 0 <= info["done_step"] <= 1001 + 255 + 1
    assert torch.isfinite(total).all()
    plan = env.tour_plan.cpu().numpy()
    for b in range(env.bs):
        seen = plan[b][plan[b] > 0]
        assert len(seen) == len(set(seen.tolist()))
    r1 = read_solomon(os.path.join(data, "R1_10_1.TXT"))
    with pytest.raises(AssertionError):
        env.load_data([r1])


def test_c1_10_1_prefix_vs_oracle():
    """BASELINE config 4 at its real size: the Gehring-Homberger file C1_10_1 (N = 1001, k = 51, K_max = 255, K6
    overrides), 2 deterministic POMO samples, the first steps of the episode teacher-forced against the CPU oracle:
    environment side bit-exact, node encoder / node embeddings / log-probabilities within the bars of the N = 301 test."""
    import os
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.lns import read_solomon
    from jampr_plus_b200.weights import random_state_dict
    from oracle.env_oracle import OracleEnv
    from oracle.policy_oracle import OraclePolicy, select_greedy
    data = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
    c1 = read_solomon(os.path.join(data, "C1_10_1.TXT"))
    kw = dict(max_concurrent_vehicles=3, k_nbh_frac=0.05, pomo=True, pomo_single_start_node=True, num_samples=2,
              tour_graph_update_step=4)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    oenv = OracleEnv(**kw)
    oenv.load([c1])
    opol = OraclePolicy(random_state_dict(0))
    env = RPEnv(inference=True, device="cuda", **kw)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
    pol.load_state_dict(random_state_dict(0))
    pol.return_logp = True
    env.load_data([c1], dist_mat=oenv.dist)
    assert env.graph_size == 1001 and env.k_nbh_size == 51 and env.max_vehicle_number == 255
    assert torch.equal(env._knn[:, 1:, :50].cpu().long(), oenv.knn[:, 1:, :50]), "kNN table at N = 1001"
    obs = env.reset()
    oenv.reset()
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore", RuntimeWarning)
        for t in range(9):              # steps 1..5 refresh the tour graph (step <= K + 1), then every 4th step
            nbh, mask = oenv.observe()
            assert torch.equal(obs.nbh.cpu().long(), nbh), f"nbh step {t}"
            assert torch.equal(obs.nbh_mask.cpu(), mask), f"mask step {t}"
            assert env._graph_fresh == oenv.graph_fresh, f"graph refresh rule step {t}"
            pol(obs)
            tour_emb, node_emb = opol.tour_encode(oenv)
            lo = opol.decode(tour_emb, node_emb, nbh, mask)
            if t == 0:
                assert (env._static_emb.cpu() - opol.static_emb).abs().max() <= EMB_ATOL, "node encoder (static_emb)"
            assert (env.state["ne"].cpu() - node_emb).abs().max() <= EMB_ATOL, f"node_emb step {t}"
            lp = pol.log_probs(env).cpu()
            fin = torch.isfinite(lo)
            assert torch.equal(fin, torch.isfinite(lp))
            assert (lp[fin] - lo[fin]).abs().max() <= LOGP_ATOL, f"log-probs step {t}"
            act, _ = select_greedy(lo, nbh)
            obs, cost, done, _ = env.step(act)
            co, _ = oenv.step(act)
            assert torch.equal(cost.cpu(), co), f"cost step {t}"
    assert torch.equal(env.tour_plan.cpu(), oenv.plan)
