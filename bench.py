#!/usr/bin/env python
"""Benchmark of the construction-rollout hot path (BASELINE.json metric: VRPTW-100 POMO rollouts/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path over one batch: per-instance tables (distance matrix, window fix-up,
kNN, depot order) + node encoder + reset with POMO start moves + the complete greedy rollout (policy step +
environment step until the batch terminates) + best-of-samples reduction.  Workload = BASELINE config 2:
4096 synthetic CVRPTW-100 instances x 16 POMO samples = 65,536 rollouts per GPU (weak scaling: every rank
gets its own 4096 instances), seeded stand-in weights (the shipped checkpoints are absent from the
reference mount).

  value  rollouts/s with the instance arrays already resident in HBM when the timed region starts
  e2e    rollouts/s through the public API with HOST (pinned) instance arrays: H2D copy of the instances and
         D2H read of the per-instance best cost + best tour plan inside the timed region
  roofline   the per-step tour-graph GNN kernel (dominant), timed with CUDA events in an instrumented episode
  cpu_baseline   the CPU oracle (a port of the reference's algorithm, oracle/) on a bounded sample

`--impl reference` times the reference's own CPU algorithm (the oracle port; the Python reference itself
cannot travel to the GPU box) on the host cores with all threads, same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRAPH_SIZE = 100
ENV_KW = dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=16, inference=True,
              tour_graph_update_step=1)
METRIC = "vrptw100_pomo_greedy_rollouts_per_sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=4096, help="instances per GPU (config 2: 4096)")
    ap.add_argument("--precision", default=os.environ.get("JAMPR_PRECISION", "fp16"), choices=["fp32", "bf16", "fp16"])
    ap.add_argument("--cpu-sample", type=int, default=12, help="instances of the CPU baseline sample (x16 samples; ~14 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def flops_tour_encoder(N, H=128, D=256):
    """Algorithmic dense FLOPs of one tour-graph GNN refresh per rollout: node_emb_input_proj +
    6 GraphConv blocks (lin_rel + lin_root) + node_emb_output_proj."""
    return 2 * N * D * H + 12 * (2 * N * H * H) + 2 * N * H * D


def flops_policy_step(N, A, K=3, H=128, D=256):
    """Algorithmic dense FLOPs of one whole policy step per rollout, F_step(N, A) of SURVEY.md §8d (the reference's
    formulation: set_proj applied to every action embedding) = 84.53 MFLOP at N=101, A=78."""
    return (flops_tour_encoder(N, H, D) + K * (2 * 5 * H + 2 * 2 * H * D) + 2 * 2 * D * D + 2 * A * D * 3 * D
            + 4 * A * D + 2 * D * D + 2 * A * D)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def make_arrays(n_inst: int, seed: int):
    from jampr_plus_b200.synth import instances_to_arrays, sample_instances
    inst = sample_instances(n_inst, graph_size=GRAPH_SIZE, seed=seed)
    arr = instances_to_arrays(inst)
    host = {k: torch.from_numpy(np.ascontiguousarray(arr[k])).to(torch.float32)
            for k in ("coords", "demands", "tw", "service_time", "org_service_horizon")}
    return inst, host, int(arr["max_vehicle_number"][0])


def cpu_port_rollouts(instances, threads: int):
    """One greedy POMO episode of the CPU oracle; returns (rollouts, seconds)."""
    from jampr_plus_b200.weights import random_state_dict
    from oracle.env_oracle import OracleEnv
    from oracle.policy_oracle import OraclePolicy, select_greedy
    torch.set_num_threads(threads)
    kw = dict(ENV_KW)
    kw.pop("inference")
    env = OracleEnv(**kw)
    pol = OraclePolicy(random_state_dict(0))
    gen = torch.Generator().manual_seed(1235)
    t0 = time.perf_counter()
    env.load(instances)
    env.reset(generator=gen)
    done = False
    with torch.no_grad():
        while not done:
            nbh, mask = env.observe()
            act, _ = select_greedy(pol.forward(env, nbh, mask), nbh)
            _, done = env.step(act)
    c = env.total.view(-1, env.P).min(-1)
    return env.bs, time.perf_counter() - t0, float(c.values.mean())


def run_reference(args, rank, world):
    if rank != 0:
        return
    from jampr_plus_b200.synth import sample_instances
    threads = os.cpu_count() or 1
    inst = sample_instances(args.cpu_sample, graph_size=GRAPH_SIZE, seed=1234)
    for _ in range(min(args.warmup, 1)):
        cpu_port_rollouts(inst[:1], threads)
    tot_r, tot_t = 0, 0.0
    for _ in range(args.steps):
        r, t, _ = cpu_port_rollouts(inst, threads)
        tot_r += r
        tot_t += t
    v = tot_r / tot_t
    sample = f"{args.cpu_sample} instances x 16 POMO samples = {args.cpu_sample * 16} rollouts per step, full episodes"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rollouts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: POMO greedy rollouts, synthetic CVRPTW-100 (all_1_low-like), 16 samples/instance",
                   "graph_size": GRAPH_SIZE + 1, "num_samples": 16, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "rollouts/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch.distributed as dist
    from jampr_plus_b200 import Policy, RPEnv, _cabi
    from jampr_plus_b200.driver import rollout, rollout_sharded
    from jampr_plus_b200.weights import random_state_dict

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    inst, host, kv = make_arrays(args.instances, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    env = RPEnv(device=dev, **ENV_KW)
    env.seed(1235 + rank)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device=dev, precision=args.precision)
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    pol.set_decode_type("greedy")
    I, P = args.instances, ENV_KW["num_samples"]
    bs = I * P
    iters = {"n": 0}

    def hot_path(arrays):
        env.load_arrays(arrays, kv, check=False)
        pol.reset_static()
        env.reset()
        if world > 1:      # union-batch termination semantics across the ranks: one 4-byte all_reduce(MAX) per episode
            rollout_sharded(pol, env)
        else:              # the whole multi-step loop (3 kernels x 129 steps) replayed as one captured CUDA graph
            rollout(pol, env, check=False, graph=True)
        iters["n"] = env.graph_size + env.max_vehicle_number + 1 - ENV_KW["max_concurrent_vehicles"] + 1
        return env.gather_best()

    out_cost = torch.empty(I * world if rank == 0 else I, dtype=torch.float32).pin_memory()
    out_plan = None

    from jampr_plus_b200.sharding import gather_best

    def e2e_step():
        nonlocal out_plan
        cost, idx, plan = hot_path(pinned)
        if world > 1:      # the one data-path collective: best tours of every shard, global instance order (NCCL)
            cost, idx, plan = gather_best(cost, idx, plan, I * world)
            if rank != 0:
                torch.cuda.current_stream().synchronize()
                return 0.0
        if out_plan is None:
            out_plan = torch.empty(plan.shape, dtype=plan.dtype).pin_memory()
        out_cost.copy_(cost, non_blocking=True)
        out_plan.copy_(plan, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(out_cost[0])

    import warnings
    warnings.simplefilter("ignore", RuntimeWarning)
    for _ in range(max(args.warmup, 3)):
        hot_path(resident)
    torch.cuda.synchronize()
    done_step = int(env.state["ctl"][_cabi.CTL_DONE_STEP].item())
    status_bits = int(env.state["status"].max().item())
    assert done_step >= 0, "episode did not terminate"
    assert os.environ.get("JAMPR_TC_ABLATE") or not (status_bits & (_cabi.ST_NO_FEASIBLE | _cabi.ST_TOUR_OVERFLOW | _cabi.ST_NAN_LOGITS)), status_bits

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # ---- timed region 1: instance arrays resident in HBM
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        hot_path(resident)
    e1.record()
    barrier()
    t_res = torch.tensor([e0.elapsed_time(e1)], device=dev)
    # ---- timed region 2: end to end from pinned host arrays, results read back to the host
    e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    t_e2e = torch.tensor([f0.elapsed_time(f1)], device=dev)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e = float(t_res.item()), float(t_e2e.item())

    # ---- instrumented episode: CUDA events around every launch of the dominant kernel (tour-graph GNN)
    roof = None
    env_roof = None
    if rank == 0:
        env.load_arrays(resident, kv, check=False)
        pol.reset_static()
        env.reset()
        st = env.state
        st.ensure_policy_buffers(False)
        w = pol.encode_static(env)
        lib, d, p, S = pol._lib, env._dims, env._inst, env._stream()
        evs, env_evs = [], []
        cap = env.graph_size + env.max_vehicle_number + 1
        step = env._step
        fused = args.precision != "fp32"      # tensor-core path: GNN + decoder are ONE kernel per step
        while step <= cap:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if fused:
                _cabi.check(lib.jampr_policy_step(d, p, w, st.struct(), 1, step, 0, 0, None, S), "jampr_policy_step")
            else:
                _cabi.check(lib.jampr_tour_encoder(d, p, w, st.struct(), step, S), "jampr_tour_encoder")
            b.record()
            evs.append((a, b))
            if not fused:
                _cabi.check(lib.jampr_policy_step(d, p, w, st.struct(), 0, step, 0, 0, None, S), "jampr_policy_step")
            c = torch.cuda.Event(enable_timing=True)
            _cabi.check(lib.jampr_env_step(d, p, st.struct(), None, step, S), "jampr_env_step")
            c.record()
            env_evs.append((b, c) if fused else None)
            step += 1
        torch.cuda.synchronize()
        done2 = int(st["ctl"][_cabi.CTL_DONE_STEP].item())
        live = [a.elapsed_time(b) for i, (a, b) in enumerate(evs) if env._step + i <= done2]
        avg_ms = float(np.mean(live))
        A_act = ENV_KW["max_concurrent_vehicles"] * env.k_nbh_size
        fl = (flops_policy_step(env.graph_size, A_act) if fused else flops_tour_encoder(env.graph_size)) * bs
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        ach = fl / (avg_ms * 1e-3) / 1e12
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "traffic_policy_step.json")
        if fused and args.precision == "fp16" and I == 4096 and os.path.exists(tr_path):
            tr = json.load(open(tr_path))      # one ncu --set full capture of this very launch shape (see `source`)
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        if fused:
            # the environment step (env_move + env_observe): HBM-bound integer / compare work, SURVEY.md §8d figure of
            # ~1.5 KB per rollout and step (candidate rows, state read + write, observation)
            live_env = [x[0].elapsed_time(x[1]) for i, x in enumerate(env_evs) if x is not None and env._step + i <= done2]
            env_ms = float(np.mean(live_env))
            A_env = ENV_KW["max_concurrent_vehicles"] * env.k_nbh_size
            env_bytes = (A_env * 14 + 16 + 140 + 3 * A_env) * bs
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            env_roof = {"bound": "hbm", "kernel": "env_move + env_observe", "achieved": env_bytes / (env_ms * 1e-3) / 1e9,
                        "peak": hbm, "unit": "GB/s", "frac": env_bytes / (env_ms * 1e-3) / 1e9 / hbm,
                        "bytes_per_launch_pair": env_bytes, "avg_ms": env_ms,
                        "share_of_step": env_ms * len(live_env) / (ms_res / args.steps)}
        roof = {"bound": "tensor", "kernel": "policy_step_tc2 (tour-graph GNN + decoder, fused)" if fused else "tour_encoder_f32",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                "flops_per_launch": fl, "avg_launch_ms": avg_ms, "launches_timed": len(live),
                "share_of_step": avg_ms * len(live) / (ms_res / args.steps)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_it = done_step - ENV_KW["max_concurrent_vehicles"] + 1          # live iterations of the device loop
    enq = env.graph_size + env.max_vehicle_number + 1 - ENV_KW["max_concurrent_vehicles"] + 1
    # own kernels per pass: setup_instances 1, edge-table prepare 1, node encoder (input, 3 GraphConv launches, projection,
    # h0: 6 on the tensor-core path, 1 in fp32), env_reset 1, per enqueued step policy 1 (fp32: 2) + env 2, gather_best 1
    tc = args.precision != "fp32"
    launches = args.steps * ((10 if tc else 4) + (3 if tc else 4) * enq)
    in_bytes = sum(v.numel() * 4 for v in pinned.values())
    out_bytes = out_cost.numel() * 4 + out_plan.numel() * 2
    total_rollouts = bs * world
    res = {
        "metric": METRIC, "value": total_rollouts * args.steps / (ms_res * 1e-3), "unit": "rollouts/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16", "fp16": "f16"}[args.precision], "data": "synthetic",
        "config": {"workload": "config2: batched POMO greedy rollouts, 4096 synthetic CVRPTW-100 instances (all_1_low-like) x 16 samples per GPU",
                   "instances_per_gpu": I, "num_samples": P, "rollouts_per_gpu": bs, "graph_size": GRAPH_SIZE + 1,
                   "weights": "seeded random init, out_proj x0.1 (shipped checkpoints absent from the reference mount)",
                   "episode_steps": done_step + 1, "device_loop_iterations_live": n_it,
                   "l2": "per-step working set (>= 10 GB of rollout state) exceeds the 126 MB L2; no explicit flush",
                   "precision": args.precision,
                   "rollout": "whole device loop replayed as one captured CUDA graph" if world == 1 else
                              "device loop enqueued per shard, union-batch termination via one all_reduce(MAX)"},
        "e2e": {"value": total_rollouts * args.steps / (ms_e2e * 1e-3), "unit": "rollouts/s",
                "h2d_bytes_per_step": in_bytes * world, "d2h_bytes_per_step": out_bytes, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof,
    }
    if env_roof is not None:
        res["roofline_env"] = env_roof
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        r, t, c = cpu_port_rollouts(inst[: args.cpu_sample], threads)
        res["cpu_baseline"] = {"value": r / t, "unit": "rollouts/s", "cores": threads, "kind": "port",
                               "sample": f"{args.cpu_sample} of the {I} instances x {P} samples = {r} rollouts, full greedy episode, oracle/ (torch CPU fp32)"}
    print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
