"""Boundary proof: the reference's OWN caller of the hot path — lib/model/training.py:54-96 `eval_episode`, imported
unmodified — runs with jampr_plus_b200.RPEnv / Policy swapped in for lib.routing.RPEnv / lib.model.Policy and returns
the same per-instance best costs as with the reference's own classes on the CPU (deterministic POMO start nodes:
`pomo_single_start_node=True`)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import ref_runner
from oracle.replay import load_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(ref_runner.reference_root() is None, reason="no copy of the reference available")]


@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-5), ("fp16", 1e-3)])
def test_reference_eval_episode_runs_on_the_mirrors(precision, rtol):
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.weights import random_state_dict
    ref = ref_runner.load_reference()
    g = load_case("g21_pomo")
    inst = g["instances"]
    kw = dict(max_concurrent_vehicles=3, k_nbh_frac=0.5, pomo=True, pomo_single_start_node=True, num_samples=6,
              inference=True, tour_graph_update_step=1)
    sd = random_state_dict(3)
    # ---- the reference, on the CPU
    RI = ref["RPInstance"]
    data_ref = [RI(**{f: getattr(x, f) for f in RI._fields}) for x in inst]
    env_r = ref["RPEnv"](device="cpu", **kw)
    env_r.seed(7)
    pol_r = ref_runner.make_reference_policy(sd)
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        cost_r, info_r = ref["training"].eval_episode(data_ref, env_r, pol_r)
        # transit matrix of the reference (MKL sqrt quirk) for the device run
        env_r.load_data(data_ref)
        dist_r = env_r._dist_mat.view(len(inst), kw["num_samples"], env_r.graph_size, env_r.graph_size)[:, 0].clone()
    # ---- the same function of the reference, driving the CUDA mirrors
    env = RPEnv(device="cuda", **kw)
    env.seed(7)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=precision)
    pol.load_state_dict(sd)
    load_data = env.load_data
    env.load_data = lambda batch: load_data(batch, dist_mat=dist_r)      # same inputs as the reference run
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cost, info = ref["training"].eval_episode(inst, env, pol)
    assert cost.shape == cost_r.shape
    if precision == "fp32":
        np.testing.assert_allclose(cost.numpy(), cost_r.numpy(), rtol=rtol)
        assert np.array_equal(np.asarray(info["k_used"]), np.asarray(info_r["k_used"]))
        assert info["max_tour_len"] == info_r["max_tour_len"]
    else:       # a near-tie may flip a tour on the 16-bit path: cost level only
        assert abs(float(cost.mean()) - float(cost_r.mean())) <= 2e-2 * float(cost_r.mean())
    assert set(info.keys()) == set(info_r.keys())
