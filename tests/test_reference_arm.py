"""The reference arm of bench.py (`--impl reference`, `cpu_baseline.kind == "reference"`) and a second pin of the
oracle: the UNMODIFIED reference (baseline/_ref, copied by baseline/make_ref.py, or /root/reference) run on the
BASELINE workload fixture, against the oracle port on the same instances and start nodes.

Skipped when no copy of the reference is present (a checkout without the build container's mount)."""
import filecmp
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import ref_runner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(ref_runner.reference_root() is None, reason="no copy of the reference available")


@pytest.mark.skipif(not (os.path.isdir("/root/reference/lib") and os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "lib"))),
                    reason="needs both the mount and the copy")
def test_copy_is_verbatim():
    mism = []
    for dp, _, fs in os.walk("/root/reference/lib"):
        if "__pycache__" in dp:
            continue
        for f in fs:
            if f.endswith(".pyc"):
                continue
            src = os.path.join(dp, f)
            dst = os.path.join(ROOT, "baseline", "_ref", os.path.relpath(src, "/root/reference"))
            if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
                mism.append(dst)
    assert not mism, mism


@needs_ref
def test_reference_and_oracle_agree_on_the_workload():
    """One greedy POMO episode of workload instance 0 (N = 101, 16 samples): the reference's own loop
    (training.py:54-96) and the oracle, started from the reference's POMO draw, end with the same tours and costs."""
    from jampr_plus_b200.synth import arrays_to_instances
    from jampr_plus_b200.weights import random_state_dict
    from oracle.env_oracle import OracleEnv
    from oracle.policy_oracle import OraclePolicy, select_greedy
    ref = ref_runner.load_reference()
    z = np.load(os.path.join(ROOT, "tests", "golden", "workload_all_1_low_4096.npz"))
    raw = {k: z[k] for k in ("coords", "demands", "tw", "service_time", "org_service_horizon", "max_vehicle_number")}
    kw = dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=16, inference=True,
              tour_graph_update_step=1)
    torch.set_num_threads(8)
    env = ref["RPEnv"](device="cpu", **kw)
    env.seed(1235)
    pol = ref_runner.make_reference_policy(random_state_dict(0))
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        pol.reset_static()
        env.load_data(ref_runner.to_ref_instances(raw, 0, 1))
        dist_ref = env._dist_mat.view(1, 16, 101, 101)[:, 0].clone()
        obs = env.reset()
        start = env.tour_plan[:, :3, 0].clone().long()
        done = False
        while not done:
            action, _, _ = pol(obs)
            obs, cost, done, info = env.step(action)
    o_env = OracleEnv(**{k: v for k, v in kw.items() if k != "inference"})
    o_env.load(arrays_to_instances({k: v[0:1] for k, v in raw.items()}), dist=dist_ref)
    o_pol = OraclePolicy(random_state_dict(0))
    o_env.reset(start_nodes=start)
    done = False
    with torch.no_grad():
        while not done:
            nbh, mask = o_env.observe()
            act, _ = select_greedy(o_pol.forward(o_env, nbh, mask), nbh)
            _, done = o_env.step(act)
    assert np.array_equal(o_env.plan.numpy(), env.tour_plan.numpy())
    np.testing.assert_allclose(o_env.total.numpy(), info["current_total_cost"], rtol=1e-6)


@needs_ref
def test_bench_reference_arm_line():
    """`bench.py --impl reference` prints one JSON line of kind "reference" (no GPU involved)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference"
    assert line["value"] > 0 and line["unit"] == "rollouts/s" and line["gpu_launches"] == 0
    assert "-like" not in line["config"]["workload"]
