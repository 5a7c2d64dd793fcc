#!/usr/bin/env python
"""LNS (construct -> destroy -> repair, no local search) on a Solomon instance for a wall-clock budget, on the GPU
path and — for the same budget — through the CPU oracle (a port of the reference's env / policy), printing one JSON
line each: iterations done, incumbent cost in original units (environment total and batch-independent tour cost).

    python tools/lns_solomon.py [--instance tests/data/c103.txt] [--time-limit 60] [--cpu-time-limit 60] [--precision fp16]

BASELINE.json configs 1 / 5.  The shipped checkpoint and OR-tools are absent, so the weights are the seeded stand-ins
and there is no local search: cost LEVELS are not comparable with the published / best-known ones (C103 826.3); the
comparable quantity is how many repair iterations fit into the budget."""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ENV_KW = dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=20, tour_graph_update_step=2)


def run_gpu(inst, args):
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.lns import LNS
    from jampr_plus_b200.weights import random_state_dict
    env = RPEnv(inference=True, device="cuda", **ENV_KW)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=args.precision)
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    lns = LNS(inst, pol, env, seed=1234)
    lns.search(num_steps=2)                                   # warm-up (library load, graph capture)
    lns = LNS(inst, pol, env, seed=1234)
    t0 = time.time()
    best, cost, true_cost = lns.search(time_limit=args.time_limit)
    return {"impl": "b200", "precision": args.precision, "seconds": time.time() - t0, "iterations": lns.step,
            "incumbent_env_cost": cost, "incumbent_tour_cost": true_cost, "n_tours": len([t for t in best if len(t) > 1]),
            "unserved": len([t for t in best if len(t) == 1]), "best_found_at_iteration": lns.best_step}


def run_cpu(inst, args):
    """Same loop through the oracle (torch CPU): selection = 2 best + 3 random, bootstrap x4, same destructor."""
    from jampr_plus_b200.lns import Destructor
    from jampr_plus_b200.weights import random_state_dict
    from oracle.env_oracle import OracleEnv
    from oracle.policy_oracle import OraclePolicy, select_greedy
    torch.set_num_threads(os.cpu_count() or 1)
    env = OracleEnv(**ENV_KW)
    pol = OraclePolicy(random_state_dict(0))
    ds = Destructor(num_partial=3, seed=1234)
    rng = np.random.default_rng(1234)
    hz = float(inst.org_service_horizon)

    def episode():
        done = False
        with torch.no_grad():
            while not done:
                nbh, mask = env.observe()
                act, _ = select_greedy(pol.forward(env, nbh, mask), nbh)
                _, done = env.step(act)
        plans = env.plan.numpy()
        full = set(range(1, env.N))
        sols = []
        for p in plans:
            tours = [r[r > 0].tolist() for r in p if (r > 0).any()]
            sols.append(tours + [[e] for e in sorted(full - {n for t in tours for n in t})])
        return sols, (env.total.numpy() * hz).tolist(), (env.true_cost().numpy() * hz).tolist()

    def select(sols, costs, tcs):
        order = np.argsort(costs, kind="stable")
        pick = order[:2].tolist() + rng.choice(order[2:], 3, replace=False).tolist()
        return [sols[i] for i in pick], [costs[i] for i in pick], [tcs[i] for i in pick]

    t0 = time.time()
    env.load([inst])
    gen = torch.Generator().manual_seed(1234)
    env.reset(generator=gen)
    sols, costs, tcs = select(*episode())
    j = int(np.argmin(costs))
    best, best_cost, best_tc, steps, slowest = sols[j], costs[j], tcs[j], 0, 0.0
    while time.time() - t0 < max(args.cpu_time_limit - slowest, 0):
        ts = time.time()
        if best_cost not in costs:
            w = int(np.argmax(costs))
            sols[w], costs[w], tcs[w] = best, best_cost, best_tc
        partial = ds.destruct([[list(t) for t in s] for s in sols] * 4)
        env.import_sol([partial])
        pol.h_cache = None if env.graph_fresh else pol.h_cache
        sols, costs, tcs = select(*episode())
        j = int(np.argmin(costs))
        if costs[j] < best_cost:
            best, best_cost, best_tc = sols[j], costs[j], tcs[j]
        steps += 1
        slowest = max(slowest, time.time() - ts)
    return {"impl": "cpu-oracle", "threads": torch.get_num_threads(), "seconds": time.time() - t0, "iterations": steps,
            "incumbent_env_cost": best_cost, "incumbent_tour_cost": best_tc}


def bench_lns(dev="cuda", precision="fp16", time_limit=20.0, cpu_time_limit=10.0, instance="c103.txt", group=None):
    """BASELINE configs 1 / 5 for bench.py (extras.lns): the all-device LNS loop (RPEnv.destruct: selection, bootstrap,
    destruction and import on tour buffers; no local search, stand-in weights) on a Solomon instance for a wall-clock
    budget; on several ranks the sub-problems are sharded (LNS.search_device docstring).  The CPU oracle port runs the
    same loop for ``cpu_time_limit`` seconds beside it (rank 0, single-GPU runs only)."""
    import torch.distributed as dist
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.lns import LNS, read_solomon
    from jampr_plus_b200.weights import random_state_dict
    warnings.simplefilter("ignore")
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    inst = read_solomon(os.path.join(ROOT, "tests", "data", instance))
    env = RPEnv(inference=True, device=dev, **ENV_KW)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device=dev, precision=precision)
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    LNS(inst, pol, env, seed=1234, selection_mode="random").search_device(num_steps=3, group=group)      # warm-up
    lns = LNS(inst, pol, env, seed=1234, selection_mode="random")
    t0 = time.time()
    plan, cost, true_cost = lns.search_device(time_limit=time_limit, group=group)
    dt = time.time() - t0
    out = {"what": f"LNS on Solomon {instance.split('.')[0]} (BASELINE config {'5' if world > 1 else '1'}): construct -> "
                   "[select -> destroy -> import -> repair]* entirely on tour buffers on the device, 20 sub-problems per "
                   "GPU and iteration, K1 overrides, NO local search (OR-tools absent), stand-in weights: cost levels are "
                   "not comparable with published ones, iterations per budget are",
           "instance": instance, "time_limit_s": time_limit, "seconds": dt, "iterations": lns.step,
           "ms_per_iteration": 1e3 * dt / max(lns.step, 1), "sub_problems_per_iteration": ENV_KW["num_samples"] * world,
           "incumbent_env_cost": cost, "incumbent_tour_cost": true_cost, "n_gpus": world,
           "n_tours": int((plan[:, 0] > 0).sum())}
    if world == 1:
        # the same loop with the device local search (relocate / 2-opt*) on every repaired solution, half the budget
        lns_ls = LNS(inst, pol, env, seed=1234, selection_mode="random", ls_step=1)
        t0 = time.time()
        _, c_ls, tc_ls = lns_ls.search_device(time_limit=0.5 * time_limit)
        dt_ls = time.time() - t0
        out["with_local_search"] = {"time_limit_s": 0.5 * time_limit, "seconds": dt_ls, "iterations": lns_ls.step,
                                    "ms_per_iteration": 1e3 * dt_ls / max(lns_ls.step, 1), "incumbent_env_cost": c_ls,
                                    "incumbent_tour_cost": tc_ls, "ls": "jampr_local_search, ls_step = 1, <= 200 moves per solution"}
    if world == 1 and cpu_time_limit > 0:
        class A:
            pass
        a = A()
        a.cpu_time_limit = cpu_time_limit
        c = run_cpu(inst, a)
        out["cpu_port"] = {"seconds": c["seconds"], "iterations": c["iterations"], "threads": c["threads"],
                           "ms_per_iteration": 1e3 * c["seconds"] / max(c["iterations"], 1),
                           "incumbent_env_cost": c["incumbent_env_cost"], "incumbent_tour_cost": c["incumbent_tour_cost"]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instance", default=os.path.join(ROOT, "tests", "data", "c103.txt"))
    ap.add_argument("--time-limit", type=float, default=60.0)
    ap.add_argument("--cpu-time-limit", type=float, default=60.0)
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    from jampr_plus_b200.lns import read_solomon
    inst = read_solomon(args.instance)
    warnings.simplefilter("ignore")
    name = os.path.basename(args.instance)
    if args.time_limit > 0:
        print(json.dumps({"instance": name, "time_limit": args.time_limit, **run_gpu(inst, args)}), flush=True)
    if args.cpu_time_limit > 0:
        print(json.dumps({"instance": name, "time_limit": args.cpu_time_limit, **run_cpu(inst, args)}), flush=True)


if __name__ == "__main__":
    main()
