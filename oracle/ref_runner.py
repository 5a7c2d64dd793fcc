"""Run the UNMODIFIED reference (jokofa/JAMPR_plus) on the CPU (TEST INFRASTRUCTURE ONLY).

Used by ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference") and by tests/test_reference_caller.py.
The reference package is imported from ``baseline/_ref`` (a verbatim copy made by baseline/make_ref.py, it travels
to the GPU box) or, in the build container, from /root/reference; its missing third-party imports are satisfied by
``oracle/ref_shims``.  Nothing under ``jampr_plus_b200/`` imports this module."""
import os
import sys
import time
import warnings
from typing import Optional

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.path.join(REPO, "baseline", "_ref"), "/root/reference"]
# config/policy/gnn_attn.yaml:5-38 of the reference, used when the yaml file itself is not next to the package
_POLICY_CFG = {
    "embedding_dim": 256,
    "node_encoder_type": "NodeEncoder",
    "node_encoder_args": {"hidden_dim": 128, "num_layers": 3, "conv_type": "GraphConv", "activation": "gelu",
                          "skip": True, "norm_type": "ln", "add_linear": False},
    "tour_encoder_type": "TourEncoder",
    "tour_encoder_args": {"hidden_dim": 128, "num_layers": 2, "propagate_reverse": True, "consolidate_nbh": True,
                          "conv_type": "GraphConv", "activation": "gelu", "skip": True, "norm_type": "ln",
                          "add_linear": False},
    "decoder_type": "AttnDecoder",
    "decoder_args": {"hidden_dim": 256, "num_heads": 4, "clip_tanh": 10.0, "bias": False},
}
_loaded = {}


def reference_root() -> Optional[str]:
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "lib", "routing")):
            return c
    return None


def load_reference():
    """Import the reference package; returns a dict of its modules / classes or raises ImportError."""
    if _loaded:
        return _loaded
    root = reference_root()
    if root is None:
        raise ImportError("the reference is not available (no baseline/_ref and no /root/reference)")
    shims = os.path.join(REPO, "oracle", "ref_shims")
    for p in (root, shims):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import lib.model.baselines  # noqa: F401  (first: breaks the reference's training <-> baselines import cycle)
        import lib.model.training as training
        from lib.model.policy import Policy
        from lib.routing import RPEnv
        from lib.routing.formats import RPInstance
    cfg = dict(_POLICY_CFG)
    yml = os.path.join(root, "config", "policy", "gnn_attn.yaml")
    if os.path.exists(yml):
        import yaml
        cfg = yaml.safe_load(open(yml))["policy_cfg"]
    _loaded.update(root=root, training=training, Policy=Policy, RPEnv=RPEnv, RPInstance=RPInstance, policy_cfg=cfg)
    return _loaded


def to_ref_instances(arrays: dict, lo: int, hi: int):
    """Instances lo..hi of a stacked workload (tests/golden/workload_*.npz layout) as the reference's RPInstance."""
    ref = load_reference()
    RI = ref["RPInstance"]
    out = []
    for i in range(lo, hi):
        n = int(arrays["coords"].shape[1])
        out.append(RI(coords=np.asarray(arrays["coords"][i], dtype=np.float32),
                      demands=np.asarray(arrays["demands"][i], dtype=np.float32),
                      tw=np.asarray(arrays["tw"][i], dtype=np.float32),
                      service_time=float(arrays["service_time"][i]),
                      graph_size=n,
                      org_service_horizon=float(arrays["org_service_horizon"][i]),
                      max_vehicle_number=int(arrays["max_vehicle_number"][i]),
                      vehicle_capacity=1.0, service_horizon=1.0, depot_idx=[0]))
    return out


def make_reference_policy(state_dict, decode: str = "greedy"):
    ref = load_reference()
    pol = ref["Policy"](ref["RPEnv"].OBSERVATION_SPACE, **ref["policy_cfg"])
    pol.load_state_dict(state_dict)
    pol.eval()
    pol.set_decode_type(decode)
    return pol


def reference_episode(instances, state_dict, env_kwargs: dict, threads: int, sampling: bool = False, seed: int = 1235):
    """One evaluation episode through the reference's own ``eval_episode`` (lib/model/training.py:54-96): returns
    (rollouts, seconds, per-instance best cost)."""
    ref = load_reference()
    torch.set_num_threads(threads)
    env = ref["RPEnv"](device="cpu", **env_kwargs)
    env.seed(seed)
    pol = make_reference_policy(state_dict)
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        costs, info = ref["training"].eval_episode(instances, env, pol, sampling=sampling)
        dt = time.perf_counter() - t0
    return len(instances) * env.num_samples, dt, costs.numpy()
