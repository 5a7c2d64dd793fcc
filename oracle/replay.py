"""Replay helpers shared by the tests (TEST INFRASTRUCTURE ONLY): load a golden case recorded from the
reference by tests/golden/make_golden.py and drive the oracle through it."""
import json
import os

import numpy as np
import torch

from jampr_plus_b200.synth import arrays_to_instances
from jampr_plus_b200.weights import random_state_dict
from .env_oracle import OracleEnv
from .policy_oracle import OraclePolicy

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CASES = ["g21_pomo", "g101_pomo", "g51_sampling", "g51_upd2", "g31_k2"]


def load_case(name: str) -> dict:
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))
    g["instances"] = arrays_to_instances({k[5:]: v for k, v in g.items() if k.startswith("inst_")})
    g["env_kwargs"] = json.loads(str(g["env_kwargs"]))
    g["decode"] = str(g["decode"])
    return g


def env_args(kw: dict) -> dict:
    return dict(max_concurrent_vehicles=kw.get("max_concurrent_vehicles", 3), k_nbh_frac=kw.get("k_nbh_frac", 0.25),
                pomo=kw.get("pomo", False), pomo_nbh_frac=kw.get("pomo_nbh_frac", 0.25),
                pomo_single_start_node=kw.get("pomo_single_start_node", False),
                num_samples=kw.get("num_samples", 1), tour_graph_update_step=kw.get("tour_graph_update_step", 1))


def start_nodes_of(g: dict):
    """Start nodes the reference drew, read back from its tour buffers after reset()."""
    kw = g["env_kwargs"]
    if not kw.get("pomo", False):
        return None
    K = kw.get("max_concurrent_vehicles", 3)
    sp = g["start_plan"]
    if kw.get("pomo_single_start_node", False):
        return torch.from_numpy(sp[:, 0, 0:1].astype(np.int64))
    return torch.from_numpy(sp[:, :K, 0].astype(np.int64))


def make_oracle(g: dict, recorded_dist: bool = True):
    """recorded_dist: replay with the transit matrix the reference itself computed (see dimacs_dist)."""
    env = OracleEnv(**env_args(g["env_kwargs"]))
    env.load(g["instances"], dist=torch.from_numpy(g["static_dist"]) if recorded_dist else None)
    pol = OraclePolicy(random_state_dict(int(g["policy_seed"])))
    return env, pol
