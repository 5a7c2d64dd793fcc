#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (`lib.routing.RPEnv`, `lib.model.Policy`) is imported from /root/reference with the
third-party import shims in oracle/ref_shims (see its README for the restated PyG / torch_cluster
semantics).  For every case the script records the instances, the POMO start nodes the reference
drew, and a per-step trace (candidate sets, feasibility masks, log-probabilities, actions, step costs)
plus the final tour buffers and total costs.  The reference has no golden vectors of its own
(SURVEY.md §4), so these traces are the pin for oracle/ and, through it, for the CUDA path.
Nothing here is read at test time except the .npz/.json files it writes.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle", "ref_shims"))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import torch  # noqa: E402
import yaml  # noqa: E402

from lib.routing import RPDataset, RPEnv  # noqa: E402  (reference)
from lib.routing.formats import RPInstance as RefInstance  # noqa: E402
from lib.model.policy import Policy  # noqa: E402  (reference)

from jampr_plus_b200.synth import sample_instances, instances_to_arrays  # noqa: E402
from jampr_plus_b200.weights import random_state_dict, state_dict_schema  # noqa: E402

torch.set_num_threads(8)
POLICY_CFG = yaml.safe_load(open(os.path.join(REF, "config/policy/gnn_attn.yaml")))["policy_cfg"]


def to_ref(instances):
    return [RefInstance(**{f: getattr(x, f) for f in RefInstance._fields}) for x in instances]


def build_policy(seed=0):
    pol = Policy(RPEnv.OBSERVATION_SPACE, **POLICY_CFG)
    ref_keys = {k: tuple(v.shape) for k, v in pol.state_dict().items()}
    mine = dict(state_dict_schema())
    assert ref_keys == mine, "state_dict schema mismatch with reference"
    assert list(pol.state_dict().keys()) == list(mine.keys()), "key order mismatch"
    pol.load_state_dict(random_state_dict(seed))
    pol.eval()
    return pol


def run_case(name, instances, env_kwargs, decode="greedy", seed=1235, policy_seed=0,
             forced_actions=None, max_steps=None):
    pol = build_policy(policy_seed)
    pol.set_decode_type(decode)
    pol.reset_static()
    env = RPEnv(inference=True, **env_kwargs)
    env.seed(seed)
    env.load_data(to_ref(instances))
    logp_box = {}
    pol.decoder.register_forward_hook(lambda m, i, o: logp_box.__setitem__("v", o.detach().clone()))
    obs = env.reset()
    K = env.max_concurrent_vehicles
    trace = {k: [] for k in ("nbh", "mask", "tour_features", "logp", "action", "ll", "cost",
                             "n_tour_edges", "total")}
    start_plan = env.tour_plan.clone().numpy()
    start_rows = env.active_to_plan_idx.clone().numpy()
    static = {
        "tw_fixed": env.tw.view(-1, env.num_samples, env.graph_size, 2)[:, 0].numpy().copy(),
        "knn": (env.nbh_edges.view(2, env.bs, -1)[0, :, env.graph_size:]
                .view(env.bs, env.graph_size - 1, -1) - env.depot_idx[:, None, None])
                .view(-1, env.num_samples, env.graph_size - 1, env.k_nbh_size)[:, 0].numpy().astype(np.int16),
        "dist": env._dist_mat.view(-1, env.num_samples, env.graph_size, env.graph_size)[:, 0].numpy().copy(),
        "k_max": np.int64(env.max_vehicle_number),
    }
    done = False
    step = 0
    with torch.no_grad():
        while not done:
            action, ll, ent = pol(obs)
            if forced_actions is not None:
                action = torch.from_numpy(forced_actions[step]).long()
            trace["nbh"].append(obs.nbh.numpy().astype(np.int16))
            trace["mask"].append(obs.nbh_mask.numpy().copy())
            trace["tour_features"].append(obs.tour_features.numpy().copy())
            trace["logp"].append(logp_box["v"].numpy().copy())
            trace["n_tour_edges"].append(np.int64(obs.tour_edges.numel() // 2 if obs.tour_edges.dim() == 2 else -1))
            trace["action"].append(action.numpy().astype(np.int16))
            trace["ll"].append(ll.numpy().copy())
            obs, cost, done, info = env.step(action)
            trace["cost"].append(cost.numpy().copy())
            trace["total"].append(env.total_cost.numpy().copy())
            step += 1
            if max_steps is not None and step >= max_steps:
                break
    if static.get("ordered_idx") is None and env.ordered_idx is not None:
        static["ordered_idx"] = env.ordered_idx.view(-1, env.num_samples, env.graph_size)[:, 0].numpy().astype(np.int16)
    out = {f"inst_{k}": v for k, v in instances_to_arrays(instances).items()}
    out.update({f"static_{k}": v for k, v in static.items()})
    out.update({f"trace_{k}": np.stack(v) for k, v in trace.items()})
    out["start_plan"] = start_plan
    out["start_rows"] = start_rows
    out["final_tour_plan"] = env.tour_plan.numpy().copy()
    out["final_total_cost"] = env.total_cost.numpy().copy()
    out["final_k_used"] = env.k_used.numpy().copy()
    out["final_visited"] = env._visited.numpy().copy()
    out["final_step"] = np.int64(env._step)
    out["env_kwargs"] = np.array(json.dumps(env_kwargs))
    out["decode"] = np.array(decode)
    out["policy_seed"] = np.int64(policy_seed)
    # the static node embedding of the first instance pins the node encoder on its own
    out["static_node_emb0"] = pol.static_node_emb[0].numpy().copy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: steps={step} bs={env.bs} final_step={env._step} "
          f"cost[:4]={out['final_total_cost'][:4]} -> {os.path.getsize(path)/1e6:.2f} MB")
    return out


def run_repair_case(name, instances, env_kwargs, seed=1235, policy_seed=0, destroy_seed=5):
    """LNS repair step (lns.py:236-275 with construct -> destroy -> import_sol -> re-rollout): a greedy POMO
    construction, a seeded destroy in the manner of env.py:_test_io (the K longest tours of every sample are cut
    to a random prefix and stay open, the rest stay complete), RPEnv.import_sol, then the recorded repair rollout."""
    pol = build_policy(policy_seed)
    pol.set_decode_type("greedy")
    pol.reset_static()
    env = RPEnv(inference=True, **env_kwargs)
    env.seed(seed)
    env.load_data(to_ref(instances))
    obs = env.reset()
    done = False
    with torch.no_grad():
        while not done:
            action, _, _ = pol(obs)
            obs, _, done, _ = env.step(action)
    prev_step = int(env._step)
    K, P = env.max_concurrent_vehicles, env.num_samples
    rng = np.random.default_rng(destroy_seed)
    plan = env.tour_plan.numpy().reshape(-1, P, *env.tour_plan.shape[1:])
    sol = []
    for inst_plan in plan:
        inst_sol = []
        for smp in inst_plan:
            tours = [r[r > 0].tolist() for r in smp if (r > 0).any()]
            order = np.argsort([-len(t) for t in tours], kind="stable")
            new = []
            for rank, ti in enumerate(order):
                t = tours[ti]
                if rank < K:
                    tlim = int(rng.integers(2, max(len(t) - 1, 3)))
                    new.append([int(x) for x in t[:tlim]])
                else:
                    new.append([0] + [int(x) for x in t] + [0])
            inst_sol.append(new)
        sol.append(inst_sol)
    # the reference needs a loaded instance for import_sol
    env.load_data(to_ref(instances))
    logp_box = {}
    pol.decoder.register_forward_hook(lambda m, i, o: logp_box.__setitem__("v", o.detach().clone()))
    obs = env.import_sol(sol)
    out = {f"inst_{k}": v for k, v in instances_to_arrays(instances).items()}
    out["static_dist"] = env._dist_mat.view(-1, P, env.graph_size, env.graph_size)[:, 0].numpy().copy()
    out["static_k_max"] = np.int64(env.max_vehicle_number)
    out["import_total"] = env.total_cost.numpy().copy()
    out["import_step"] = np.int64(env._step)
    out["import_plan"] = env.tour_plan.numpy().copy()
    out["import_rows"] = env.active_to_plan_idx.numpy().copy()
    out["import_visited"] = env._visited.numpy().copy()
    out["prev_step"] = np.int64(prev_step)
    out["solution_json"] = np.array(json.dumps(sol))
    trace = {k: [] for k in ("nbh", "mask", "tour_features", "logp", "action", "cost", "n_tour_edges", "total")}
    done = False
    with torch.no_grad():
        while not done:
            action, ll, ent = pol(obs)
            trace["nbh"].append(obs.nbh.numpy().astype(np.int16))
            trace["mask"].append(obs.nbh_mask.numpy().copy())
            trace["tour_features"].append(obs.tour_features.numpy().copy())
            trace["logp"].append(logp_box["v"].numpy().copy())
            trace["n_tour_edges"].append(np.int64(obs.tour_edges.numel() // 2 if obs.tour_edges.dim() == 2 else -1))
            trace["action"].append(action.numpy().astype(np.int16))
            obs, cost, done, info = env.step(action)
            trace["cost"].append(cost.numpy().copy())
            trace["total"].append(env.total_cost.numpy().copy())
    out.update({f"trace_{k}": np.stack(v) for k, v in trace.items()})
    out["final_tour_plan"] = env.tour_plan.numpy().copy()
    out["final_total_cost"] = env.total_cost.numpy().copy()
    out["final_visited"] = env._visited.numpy().copy()
    out["final_step"] = np.int64(env._step)
    out["env_kwargs"] = np.array(json.dumps(env_kwargs))
    out["decode"] = np.array("greedy")
    out["policy_seed"] = np.int64(policy_seed)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: construct ended at step {prev_step}; import -> step {int(out['import_step'])}, "
          f"repair steps={len(trace['action'])}, final_step={env._step} -> {os.path.getsize(path)/1e6:.2f} MB")
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "repair":
        run_repair_case("g51_repair", sample_instances(3, graph_size=50, seed=21),
                        dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=6))
        return
    with open(os.path.join(HERE, "state_dict_schema.json"), "w") as f:
        json.dump({k: list(v) for k, v in state_dict_schema().items()}, f, indent=1)

    # case 1: tiny graphs, my generator, POMO with K random start nodes per rollout
    run_case("g21_pomo", sample_instances(6, graph_size=20, seed=7),
             dict(max_concurrent_vehicles=3, k_nbh_frac=0.5, pomo=True, pomo_nbh_frac=0.5, num_samples=4))

    # case 2: VRPTW-100, instances drawn by the REFERENCE generator (all_1_low), POMO greedy
    ds = RPDataset(cfg={"groups": ["r", "c", "rc"], "types": [1], "tw_fracs": [0.25, 0.5]},
                   stats_pth=os.path.join(REF, "solomon_stats.pkl"))
    ds.seed(1234)
    ref_inst = ds.sample(sample_size=8, graph_size=100).data
    pick = [ref_inst[0], ref_inst[3], ref_inst[5]]   # one r, one c, one rc sampler
    run_case("g101_pomo", pick,
             dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=8))

    # case 3: sampling decode without POMO (inference num_samples), N=51, type-2-like neighbourhood
    run_case("g51_sampling", sample_instances(4, graph_size=50, seed=11, types=(2,), tw_fracs=(0.75, 1.0)),
             dict(max_concurrent_vehicles=3, k_nbh_frac=0.32, pomo=False, num_samples=4),
             decode="sampling", seed=99)

    # case 4: single POMO start node + tour graph refreshed every 2nd step (LNS K1 override)
    run_case("g51_upd2", sample_instances(3, graph_size=50, seed=13),
             dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, pomo_single_start_node=True,
                  num_samples=6, tour_graph_update_step=2))

    # case 5: a second weight seed and 2 concurrent vehicles
    run_case("g31_k2", sample_instances(4, graph_size=30, seed=17),
             dict(max_concurrent_vehicles=2, k_nbh_frac=0.4, pomo=True, pomo_nbh_frac=0.5, num_samples=4),
             policy_seed=3)

    # case 6: the LNS repair step: construct -> destroy -> import_sol -> repair rollout
    run_repair_case("g51_repair", sample_instances(3, graph_size=50, seed=21),
                    dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=6))


if __name__ == "__main__":
    main()
