#!/usr/bin/env python
"""BASELINE config 4 shape: 1000-customer instances (N = 1001, K6 overrides: k_nbh_frac 0.05, tour_graph_update_step 4),
full-length greedy POMO rollouts through the large-graph path.  Prints one JSON line: seconds per episode, rollouts/s,
steps, peak device memory; optionally times the CPU oracle on a step prefix (the reference itself needed 28.9 s for
the first 200 steps of ONE instance x 4 samples in the survey probe).

    python tools/bench_gh1000.py [--instances 32] [--samples 4] [--cpu-steps 0]"""
import argparse
import json
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bench_gh1000(dev="cuda", precision="fp16", instances=60, samples=4, reps=2, world=1):
    """BASELINE config 4 shape on one GPU (every rank of a multi-GPU run executes this on its own batch: weak scaling);
    returns the dict bench.py prints under extras.gh1000."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.synth import sample_instances
    from jampr_plus_b200.weights import random_state_dict
    warnings.simplefilter("ignore")
    torch.cuda.reset_peak_memory_stats(dev)
    base = torch.cuda.memory_allocated(dev)
    # Gehring-Homberger 1000-customer instances allow 250 vehicles (K_max = 255, step cap N + K_max + 1 = 1257)
    inst = [x._replace(max_vehicle_number=250) for x in sample_instances(instances, graph_size=1000, seed=9)]
    kw = dict(max_concurrent_vehicles=3, k_nbh_frac=0.05, pomo=True, num_samples=samples, tour_graph_update_step=4)
    env = RPEnv(inference=True, device=dev, **kw)
    env.seed(1)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device=dev, precision=precision)
    pol.load_state_dict(random_state_dict(0))
    times = []
    for it in range(1 + reps):
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.load_data(inst)
        pol.reset_static()
        env.reset()
        total, info = rollout(pol, env)
        cost, idx, plan = env.gather_best()
        cost = cost.cpu()
        e1.record()
        torch.cuda.synchronize(dev)
        times.append(e0.elapsed_time(e1) * 1e-3)
    sec = min(times[1:])
    return {"what": "BASELINE config 4 shape: 1000-customer instances (N = 1001, k = 51, K_max = %d, K6 overrides: "
                    "k_nbh_frac 0.05, tour_graph_update_step 4), full-length greedy POMO episodes, layer-wise tcgen05 "
                    "GNN path" % env.max_vehicle_number,
            "instances_per_gpu": instances, "samples": samples, "rollouts_per_gpu": env.bs,
            "episode_steps": info["done_step"] + 1, "seconds_per_episode": sec, "value": env.bs * world / sec,
            "unit": "rollouts/s", "policy_steps_per_s": env.bs * world * (info["done_step"] + 1) / sec,
            "peak_mem_gb_per_gpu": (torch.cuda.max_memory_allocated(dev) - base) / 2**30,
            "mean_best_cost": float(cost.mean()), "precision": precision, "n_gpus": world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=32)
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--cpu-steps", type=int, default=0)
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.synth import sample_instances
    from jampr_plus_b200.weights import random_state_dict
    warnings.simplefilter("ignore")
    # Gehring-Homberger 1000-customer instances allow 250 vehicles (K_max = 255, step cap N + K_max + 1 = 1257)
    inst = [x._replace(max_vehicle_number=250) for x in sample_instances(args.instances, graph_size=1000, seed=9)]
    kw = dict(max_concurrent_vehicles=3, k_nbh_frac=0.05, pomo=True, num_samples=args.samples, tour_graph_update_step=4)
    env = RPEnv(inference=True, device="cuda", **kw)
    env.seed(1)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=args.precision)
    pol.load_state_dict(random_state_dict(0))
    times = []
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        env.load_data(inst)
        pol.reset_static()
        env.reset()
        total, info = rollout(pol, env)
        cost, idx, plan = env.gather_best()
        cost = cost.cpu()
        torch.cuda.synchronize()
        times.append(time.time() - t0)
    sec = min(times[1:])
    out = {"workload": "config4 shape: N=1001, k=51, K_max=%d" % env.max_vehicle_number, "instances": args.instances,
           "samples": args.samples, "rollouts": env.bs, "episode_steps": info["done_step"] + 1, "seconds_per_episode": sec,
           "rollouts_per_s": env.bs / sec, "policy_steps_per_s": env.bs * (info["done_step"] + 1) / sec,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, "mean_best_cost": float(cost.mean()),
           "precision": args.precision}
    if args.cpu_steps > 0:
        from oracle.env_oracle import OracleEnv
        from oracle.policy_oracle import OraclePolicy, select_greedy
        torch.set_num_threads(os.cpu_count() or 1)
        oenv = OracleEnv(**kw)
        t0 = time.time()
        oenv.load(inst[:1])
        opol = OraclePolicy(random_state_dict(0))
        oenv.reset(generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            for _ in range(args.cpu_steps):
                nbh, mask = oenv.observe()
                act, _ = select_greedy(opol.forward(oenv, nbh, mask), nbh)
                oenv.step(act)
        dt = time.time() - t0
        out["cpu_oracle"] = {"instances": 1, "samples": args.samples, "steps": args.cpu_steps, "seconds": dt,
                             "policy_steps_per_s": args.samples * args.cpu_steps / dt, "threads": torch.get_num_threads()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
