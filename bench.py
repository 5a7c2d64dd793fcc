#!/usr/bin/env python
"""Benchmark of the construction-rollout hot path (BASELINE.json metric: VRPTW-100 POMO rollouts/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp16|fp32|bf16]

A *step* is one pass of the hot path over one batch: per-instance tables (distance matrix, window fix-up, kNN, depot
order) + node encoder + reset with POMO start moves + the complete greedy rollout (policy step + environment step until
the batch terminates) + best-of-samples reduction.

Workload = BASELINE config 2: the 4096 CVRPTW-100 instances `RPDataset(all_1_low).seed(1234).sample(4096, 100)` of the
reference generator (tests/golden/workload_all_1_low_4096.npz, written by tests/golden/make_workload.py from the
unmodified reference) x 16 POMO samples = 65,536 rollouts per GPU; weak scaling: every rank runs the same 4096
instances with its own POMO start draw.  Seeded stand-in weights (the shipped checkpoints are absent from the reference
mount).

  value      rollouts/s with the instance arrays already resident in HBM when the timed region starts
  e2e        rollouts/s through the public API with HOST (pinned) instance arrays: H2D copy of the instances and D2H
             read of the per-instance best cost + best tour plan inside the timed region (N > 1: + the NCCL gather)
  roofline   the dominant kernel (fused policy step), CUDA events around every launch of an instrumented episode;
             roofline.secondary = the environment step against the HBM roofline
  parity     log-probability error of this very build / precision against the reference trace g101_pomo (teacher forced)
  cpu_baseline   the unmodified reference (baseline/_ref, kind "reference") or, when it is absent, the oracle port, on a
             bounded sample on the host cores
  extras     the other BASELINE configs measured in the same run: strong scaling (4096 instances split over the ranks),
             config 3 (sampling, 128 samples x 512 all_2_high instances, sharded), fp32 / bf16 values, GH-1000, LNS

`--impl reference` times the reference's own CPU implementation (lib.routing.RPEnv + lib.model.Policy through
lib.model.training.eval_episode, unmodified, from baseline/_ref) on the host cores with all threads, same metric.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRAPH_SIZE = 100
ENV_KW = dict(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=16, inference=True,
              tour_graph_update_step=1)
ENV_KW_C3 = dict(max_concurrent_vehicles=3, k_nbh_frac=0.32, pomo=False, num_samples=128, inference=True,
                 tour_graph_update_step=1)          # config/env/all_2_high.yaml:15, SURVEY.md §8d C3
METRIC = "vrptw100_pomo_greedy_rollouts_per_sec"
WL_C2 = os.path.join(ROOT, "tests", "golden", "workload_all_1_low_4096.npz")
WL_C3 = os.path.join(ROOT, "tests", "golden", "workload_all_2_high_512.npz")
WORKLOAD_NAME = ("config2: batched POMO greedy rollouts, RPDataset(all_1_low).seed(1234).sample(4096, graph_size=100) "
                 "of the reference generator x 16 POMO samples per GPU")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=4096, help="instances per GPU (config 2: 4096)")
    ap.add_argument("--precision", default=os.environ.get("JAMPR_PRECISION", "fp16"), choices=["fp32", "bf16", "fp16"])
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="instances per step of the CPU arm (x16 samples); 0 = max(1, ceil(12 / steps))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default="all", help="comma list of strong,config3,fp32,bf16,gh1000,lns | all | none")
    ap.add_argument("--lns-time", type=float, default=20.0, help="wall-clock budget of the LNS extra (BASELINE metric: 60)")
    ap.add_argument("--no-status-check", action="store_true", help="ablation builds only: skip the status-word assertion")
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


def load_workload(path: str, lo: int = 0, hi: int = None):
    """(raw numpy arrays, host fp32 tensors of instances lo..hi, max_vehicle_number)."""
    z = np.load(path)
    hi = int(z["coords"].shape[0]) if hi is None else hi
    raw = {k: z[k] for k in ("coords", "demands", "tw", "service_time", "org_service_horizon", "max_vehicle_number")}
    host = {k: torch.from_numpy(np.ascontiguousarray(raw[k][lo:hi])).to(torch.float32)
            for k in ("coords", "demands", "tw", "service_time", "org_service_horizon")}
    return raw, host, int(raw["max_vehicle_number"][0])


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_port_rollouts(raw, lo, hi, threads: int):
    """One greedy POMO episode of the CPU oracle (a port of the reference's algorithm); (rollouts, seconds)."""
    from jampr_plus_b200.synth import arrays_to_instances
    from jampr_plus_b200.weights import random_state_dict
    from oracle.env_oracle import OracleEnv
    from oracle.policy_oracle import OraclePolicy, select_greedy
    torch.set_num_threads(threads)
    instances = arrays_to_instances({k: v[lo:hi] for k, v in raw.items()})
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
    return env.bs, time.perf_counter() - t0


def cpu_arm(raw, lo, hi, threads):
    """The reference itself when its copy is present (baseline/_ref or /root/reference), else the oracle port.
    Returns (kind, rollouts, seconds, description)."""
    from jampr_plus_b200.weights import random_state_dict
    try:
        from oracle import ref_runner
        ref_runner.load_reference()
    except Exception as e:            # noqa: BLE001 - any import problem of the reference copy means "not available"
        r, t = cpu_port_rollouts(raw, lo, hi, threads)
        return "port", r, t, f"oracle/ port (reference copy unavailable: {type(e).__name__}: {e})"
    inst = ref_runner.to_ref_instances(raw, lo, hi)
    r, t, _ = ref_runner.reference_episode(inst, random_state_dict(0), ENV_KW, threads)
    return "reference", r, t, ("unmodified reference lib.routing.RPEnv + lib.model.Policy via lib.model.training.eval_episode "
                               "(baseline/_ref; scatter-max / kNN through the pure-torch shims of oracle/ref_shims)")


def run_reference(args, rank, world):
    if rank != 0:
        return
    raw, _, _ = load_workload(WL_C2)
    threads = os.cpu_count() or 1
    per_step = args.cpu_sample if args.cpu_sample > 0 else max(1, math.ceil(12 / max(args.steps, 1)))
    n_total = raw["coords"].shape[0]
    kind = desc = None
    for _ in range(min(args.warmup, 1)):
        cpu_arm(raw, 0, 1, threads)
    tot_r, tot_t, lo = 0, 0.0, 0
    for _ in range(args.steps):
        hi = lo + per_step
        if hi > n_total:
            lo, hi = 0, per_step
        kind, r, t, desc = cpu_arm(raw, lo, hi, threads)
        tot_r += r
        tot_t += t
        lo = hi
    v = tot_r / tot_t
    sample = (f"{per_step} instance(s) x 16 POMO samples = {per_step * 16} rollouts per step, a fresh slice of the 4096 "
              f"workload instances every step ({tot_r} rollouts in total), full greedy episodes; {desc}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rollouts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "graph_size": GRAPH_SIZE + 1, "num_samples": 16, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "rollouts/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm helpers
def parity_probe(dev, precision):
    """Teacher-forced log-probability error of this build against the reference trace g101_pomo (tests/golden): the
    figures tests/test_gpu_tc_policy.py asserts on, measured in the bench process."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.synth import arrays_to_instances
    from jampr_plus_b200.weights import random_state_dict
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "g101_pomo.npz"), allow_pickle=False))
    kw = json.loads(str(g["env_kwargs"]))
    inst = arrays_to_instances({k[5:]: v for k, v in g.items() if k.startswith("inst_")})
    env = RPEnv(inference=True, device=dev, max_concurrent_vehicles=kw.get("max_concurrent_vehicles", 3),
                k_nbh_frac=kw.get("k_nbh_frac", 0.25), pomo=kw.get("pomo", False),
                pomo_nbh_frac=kw.get("pomo_nbh_frac", 0.25), num_samples=kw.get("num_samples", 1),
                tour_graph_update_step=kw.get("tour_graph_update_step", 1))
    pol = Policy(RPEnv.OBSERVATION_SPACE, device=dev, precision=precision, allow_reduced_accuracy=True)
    pol.load_state_dict(random_state_dict(int(g["policy_seed"])))
    pol.eval()
    pol.return_logp = True
    env.load_data(inst, dist_mat=torch.from_numpy(g["static_dist"]))
    K = kw.get("max_concurrent_vehicles", 3)
    sn = torch.from_numpy(g["start_plan"][:, :K, 0].astype(np.int32))
    obs = env.reset(start_nodes=sn)
    T = g["trace_action"].shape[0]
    abs_max = rel_max = 0.0
    flips = flips_gap = decisions = 0
    for t in range(T):
        pol(obs)
        lp = pol.log_probs(env).cpu().numpy()
        ref = g["trace_logp"][t]
        fin = np.isfinite(ref)
        err = np.abs(lp - ref)[fin]
        abs_max = max(abs_max, float(err.max()))
        rel_max = max(rel_max, float((err / np.maximum(1.0, np.abs(ref[fin]))).max()))
        srt = np.sort(ref, -1)
        gap = srt[:, -1] - srt[:, -2]
        same = (env.state["action"].cpu().numpy() == g["trace_action"][t]).all(-1)
        flips += int((~same).sum())
        flips_gap += int((~same & (gap > 1e-3)).sum())
        decisions += same.size
        obs, _, done, _ = env.step(torch.from_numpy(g["trace_action"][t].astype(np.int64)))
    env_exact = bool(np.array_equal(env.tour_plan.cpu().numpy(), g["final_tour_plan"]))
    return {"case": "g101_pomo (reference trace, teacher forced, 24 rollouts x 128 steps)", "precision": precision,
            "logp_abs_max": abs_max, "logp_rel_max": rel_max, "rel_definition": "|dlogp| / max(1, |logp_ref|)",
            "decisions": decisions, "decisions_flipped": flips, "decisions_flipped_gap>1e-3": flips_gap,
            "env_bit_exact": env_exact}


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
    from jampr_plus_b200.sharding import gather_best, shard_bounds
    from jampr_plus_b200.weights import random_state_dict

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    warnings.simplefilter("ignore", RuntimeWarning)
    extras_sel = set() if args.extras == "none" else \
        ({"strong", "config3", "fp32", "bf16", "gh1000", "lns"} if args.extras == "all" else set(args.extras.split(",")))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """steps calls of fn bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    raw, host, kv = load_workload(WL_C2, 0, args.instances)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    env = RPEnv(device=dev, **ENV_KW)
    env.seed(1235 + rank)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device=dev, precision=args.precision, allow_reduced_accuracy=True)
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    pol.set_decode_type("greedy")
    I, P = args.instances, ENV_KW["num_samples"]
    bs = I * P

    def episode(e, p, arrays, kv_, sharded):
        e.load_arrays(arrays, kv_, check=False)
        p.reset_static()
        e.reset()
        if sharded:        # union-batch termination semantics across the ranks: one 4-byte all_reduce(MAX) per episode;
            rollout_sharded(p, e, graph=True)      # the device loop up to the local finishing step is one captured graph
        else:              # the whole multi-step loop (3 kernels x 129 steps) replayed as one captured CUDA graph
            rollout(p, e, check=False, graph=True)
        return e.gather_best()

    def hot_path(arrays):
        return episode(env, pol, arrays, kv, world > 1)

    out_cost = torch.empty(I * world if rank == 0 else I, dtype=torch.float32).pin_memory()
    out_plan = None

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

    for _ in range(max(args.warmup, 3)):
        hot_path(resident)
    torch.cuda.synchronize()
    done_step = int(env.state["ctl"][_cabi.CTL_DONE_STEP].item())
    status_bits = int(env.state["status"].max().item())
    assert done_step >= 0, "episode did not terminate"
    if not args.no_status_check:
        assert not (status_bits & (_cabi.ST_NO_FEASIBLE | _cabi.ST_TOUR_OVERFLOW | _cabi.ST_NAN_LOGITS)), status_bits
    mean_best = float(env.gather_best()[0].mean().item())

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_res = timed(lambda: hot_path(resident), args.steps)       # timed region 1: instance arrays resident in HBM
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)                         # timed region 2: end to end from pinned host arrays
    clk = clocks.stop() if rank == 0 else None

    extras = {}
    # ---- strong scaling: the SAME 4096 instances split over the ranks (contiguous blocks), gathered on rank 0
    if "strong" in extras_sel and world > 1:
        lo, hi = shard_bounds(I, rank, world)
        sh_pinned = {k: v[lo:hi].contiguous().pin_memory() for k, v in host.items()}
        env_s = RPEnv(device=dev, **ENV_KW)
        env_s.seed(2235 + rank)
        s_cost = torch.empty(I, dtype=torch.float32).pin_memory()

        def strong_step():
            cost, idx, plan = episode(env_s, pol, sh_pinned, kv, True)
            cost, idx, plan = gather_best(cost, idx, plan, I)
            if rank == 0:
                s_cost.copy_(cost, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            strong_step()
        ks = max(2, min(args.steps, 5))
        ms = timed(strong_step, ks)
        extras["strong"] = {"value": bs * ks / (ms * 1e-3), "unit": "rollouts/s", "ms_per_step": ms / ks, "steps": ks,
                            "scaling": "strong", "instances_total": I, "instances_per_gpu": hi - lo,
                            "what": "config 2 with the 4096 instances split over the ranks, end to end from pinned host "
                                    "arrays incl. the NCCL gather of the best tours"}
        del env_s
    # ---- config 3: sampling rollouts, 128 samples x 512 all_2_high instances, instances sharded over the ranks
    if "config3" in extras_sel and os.path.exists(WL_C3):
        raw3, host3, kv3 = load_workload(WL_C3)
        I3 = host3["coords"].shape[0]
        lo, hi = shard_bounds(I3, rank, world)
        pin3 = {k: v[lo:hi].contiguous().pin_memory() for k, v in host3.items()}
        env3 = RPEnv(device=dev, **ENV_KW_C3)
        env3.seed(3235 + rank)
        pol3 = Policy(RPEnv.OBSERVATION_SPACE, device=dev, precision=args.precision, allow_reduced_accuracy=True)
        pol3.load_state_dict(random_state_dict(0))
        pol3.eval()
        pol3.set_decode_type("sampling")
        pol3.sampling_seed = 20260101 + rank
        c3_cost = torch.empty(I3, dtype=torch.float32).pin_memory()

        def c3_step():
            cost, idx, plan = episode(env3, pol3, pin3, kv3, world > 1)
            if world > 1:
                cost, idx, plan = gather_best(cost, idx, plan, I3)
            if rank == 0:
                c3_cost.copy_(cost, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            c3_step()
        k3 = max(2, min(args.steps, 3))
        ms = timed(c3_step, k3)
        extras["config3"] = {"value": I3 * 128 * k3 / (ms * 1e-3), "unit": "rollouts/s", "ms_per_step": ms / k3, "steps": k3,
                             "scaling": "strong" if world > 1 else "single", "instances_total": I3, "num_samples": 128,
                             "k_nbh_frac": 0.32, "decode": "sampling (counter-based in-kernel generator, fixed seed)",
                             "mean_best_cost": float(c3_cost.mean()) if rank == 0 else None,
                             "what": "BASELINE config 3: RPDataset(all_2_high).seed(1234).sample(512, 100) x 128 sampled "
                                     "rollouts, end to end from pinned host arrays"}
        del env3, pol3

    # ---- instrumented episode: CUDA events around every launch of the dominant kernel
    roof = None
    env_roof = None
    par = None
    if rank == 0:
        env.load_arrays(resident, kv, check=False)
        pol.reset_static()
        env.reset()
        st = env.state
        st.ensure_policy_buffers(False, split=args.precision != "fp32")
        w = pol.encode_static(env)
        lib, d, p, S = pol._lib, env._dims, env._inst, env._stream()
        evs, env_evs = [], []
        cap = env.graph_size + env.max_vehicle_number + 1
        step = env._step
        fused = args.precision != "fp32"      # tensor-core path: GNN + decoder are ONE kernel per step
        split = fused and bs >= _cabi.SPLIT_MIN_ROLLOUTS and pol.tc_variant in (0, 3)
        dec_evs = []
        while step <= cap:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if split:          # the two launches of the policy step one by one: GNN kernel, then the 128-rollout decoder
                m = torch.cuda.Event(enable_timing=True)
                _cabi.check(lib.jampr_policy_step_part(d, p, w, st.struct(), 1, step, 0, 0, None, 1, S), "jampr_policy_step_part")
                m.record()
                _cabi.check(lib.jampr_policy_step_part(d, p, w, st.struct(), 1, step, 0, 0, None, 2, S), "jampr_policy_step_part")
                dec_evs.append(m)
            elif fused:
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
            env_bytes = (A_act * 14 + 16 + 140 + 3 * A_act) * bs
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            env_roof = {"bound": "hbm", "kernel": "env_move + env_observe", "achieved": env_bytes / (env_ms * 1e-3) / 1e9,
                        "peak": hbm, "unit": "GB/s", "frac": env_bytes / (env_ms * 1e-3) / 1e9 / hbm, "traffic": None,
                        "bytes_per_launch_pair": env_bytes, "avg_ms": env_ms,
                        "share_of_step": env_ms * len(live_env) / (ms_res / args.steps)}
        roof = {"bound": "tensor", "kernel": "policy step = policy_step_tc2 (tour-graph GNN, split mode) + decoder_tc (tcgen05 "
                                             "decoder over 128 rollouts), both launches timed together" if fused else "tour_encoder_f32",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                "flops_per_launch": fl, "avg_launch_ms": avg_ms, "launches_timed": len(live),
                "share_of_step": avg_ms * len(live) / (ms_res / args.steps)}
        if env_roof is not None:
            roof["secondary"] = env_roof
        if dec_evs:
            # the decoder launch against the HBM roofline.  Algorithmic bytes per rollout and step = what the launch must
            # read and write whatever its organisation (decoder tail "fused behind set_proj"): the candidate
            # rows A*D*s once, the node statistics 2*D*4, the per-tour means and features K*(128+8)*4, candidate ids and
            # masks 3*A, the decision 24 B.  Its hand-over vectors between stages (q, tour embeddings, qp, gbar, lp:
            # ~40 KB per rollout written and re-read, mostly through L2) are implementation traffic: see `traffic`.
            live_dec = [m.elapsed_time(b) for i, (m, (_, b)) in enumerate(zip(dec_evs, evs)) if env._step + i <= done2]
            live_gnn = [a.elapsed_time(m) for i, (m, (a, _)) in enumerate(zip(dec_evs, evs)) if env._step + i <= done2]
            dec_ms = float(np.mean(live_dec))
            Kv = ENV_KW["max_concurrent_vehicles"]
            dec_bytes = (A_act * 256 * 2 + 2 * 256 * 4 + Kv * (128 + 8) * 4 + 3 * A_act + 24) * bs
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            dtr = tr.get("per_kernel", {}).get("decoder_tc_kernel") if traffic is not None else None
            roof["decoder"] = {"bound": "hbm", "kernel": "decoder_tc_kernel (second launch of the policy step)",
                               "achieved": dec_bytes / (dec_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                               "frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / hbm,
                               "traffic": None if not dtr else dtr["dram__bytes_read.sum"] + dtr["dram__bytes_write.sum"],
                               "bytes_per_launch": dec_bytes, "avg_ms": dec_ms,
                               "share_of_step": dec_ms * len(live_dec) / (ms_res / args.steps)}
            roof["gnn_kernel_avg_ms"] = float(np.mean(live_gnn))
        par = parity_probe(dev, args.precision)

    # ---- BASELINE config 4 (GH-1000 shape, weak scaling) and configs 1 / 5 (LNS; sharded sub-problems for N > 1): all ranks
    del resident
    torch.cuda.empty_cache()
    if "gh1000" in extras_sel:
        try:
            from tools.bench_gh1000 import bench_gh1000
            barrier()
            r = bench_gh1000(dev, precision=args.precision, world=world)
            if world > 1:
                t = torch.tensor([r["seconds_per_episode"]], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                r["seconds_per_episode"] = float(t.item())
                r["value"] = r["rollouts_per_gpu"] * world / r["seconds_per_episode"]
                r["scaling"] = "weak"
            extras["gh1000"] = r
        except Exception as e:      # noqa: BLE001 - an extra must never take the headline line down
            if world > 1:
                raise
            extras["gh1000"] = {"error": f"{type(e).__name__}: {e}"}
    if "lns" in extras_sel:
        try:
            from tools.lns_solomon import bench_lns
            barrier()
            extras["lns"] = bench_lns(dev, precision=args.precision, time_limit=args.lns_time,
                                      instance="c103.txt" if world == 1 else "r101.txt")
        except Exception as e:      # noqa: BLE001
            if world > 1:
                raise
            extras["lns"] = {"error": f"{type(e).__name__}: {e}"}
    resident = {k: v.to(dev) for k, v in host.items()}

    # ---- other precisions on rank 0 (N = 1 runs only): same workload, resident arrays
    if rank == 0 and world == 1:
        for prec, n_i in (("fp32", 512), ("bf16", I)):
            if prec not in extras_sel or prec == args.precision:
                continue
            env_p = RPEnv(device=dev, **ENV_KW)
            env_p.seed(1235)
            pol_p = Policy(RPEnv.OBSERVATION_SPACE, device=dev, precision=prec, allow_reduced_accuracy=True)
            pol_p.load_state_dict(random_state_dict(0))
            pol_p.eval()
            arr = {k: v[:n_i].contiguous() for k, v in resident.items()}
            for _ in range(2):
                episode(env_p, pol_p, arr, kv, False)
            kp = 2
            ms = timed(lambda: episode(env_p, pol_p, arr, kv, False), kp)
            extras[prec] = {"value": n_i * P * kp / (ms * 1e-3), "unit": "rollouts/s", "ms_per_step": ms / kp,
                            "instances": n_i, "what": f"config 2 on the {prec} path, first {n_i} instances, resident arrays" +
                            (" (the 1e-5 parity anchor, CUDA cores)" if prec == "fp32" else
                             " (reduced accuracy: not a supported product precision)")}
            del env_p, pol_p
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_it = done_step - ENV_KW["max_concurrent_vehicles"] + 1          # live iterations of the device loop
    enq = env.graph_size + env.max_vehicle_number + 1 - ENV_KW["max_concurrent_vehicles"] + 1
    # own kernels per pass: setup_instances 1, POMO start nodes 1, edge-table prepare 1, node encoder (input, 3 GraphConv
    # launches, projection, h0: 6 on the tensor-core path, 1 in fp32), env_reset 1, per enqueued step policy (GNN kernel +
    # 128-rollout decoder: 2; fused variant 1; fp32: tour encoder + decoder 2) + env 2, gather_best 1
    tc = args.precision != "fp32"
    split_run = tc and bs >= _cabi.SPLIT_MIN_ROLLOUTS and pol.tc_variant in (0, 3)
    per_step = (4 if split_run else 3) if tc else 4
    launches = args.steps * ((11 if tc else 5) + per_step * enq)
    in_bytes = sum(v.numel() * 4 for v in pinned.values())
    out_bytes = out_cost.numel() * 4 + out_plan.numel() * 2
    total_rollouts = bs * world
    res = {
        "metric": METRIC, "value": total_rollouts * args.steps / (ms_res * 1e-3), "unit": "rollouts/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16", "fp16": "f16"}[args.precision], "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME,
                   "instances_per_gpu": I, "num_samples": P, "rollouts_per_gpu": bs, "graph_size": GRAPH_SIZE + 1,
                   "weights": "seeded random init, out_proj x0.1 (shipped checkpoints absent from the reference mount)",
                   "episode_steps": done_step + 1, "device_loop_iterations_live": n_it,
                   "mean_best_cost": mean_best,
                   "l2": "per-step working set (>= 10 GB of rollout state) exceeds the 126 MB L2; no explicit flush",
                   "precision": args.precision,
                   "rollout": "whole device loop replayed as one captured CUDA graph" if world == 1 else
                              "device loop up to the shard's finishing step replayed as one captured CUDA graph, "
                              "union-batch termination via one all_reduce(MAX)"},
        "e2e": {"value": total_rollouts * args.steps / (ms_e2e * 1e-3), "unit": "rollouts/s",
                "h2d_bytes_per_step": in_bytes * world, "d2h_bytes_per_step": out_bytes, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof,
        "parity": par,
        "extras": extras,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_cpu = args.cpu_sample if args.cpu_sample > 0 else 6      # ~10-20 s of the reference on the host cores
        kind, r, t, desc = cpu_arm(raw, 0, n_cpu, threads)
        if kind == "port" and args.cpu_sample <= 0:      # the port is ~5x faster than the reference: a larger sample
            kind, r, t, desc = cpu_arm(raw, 0, 12, threads)
        res["cpu_baseline"] = {"value": r / t, "unit": "rollouts/s", "cores": threads, "kind": kind,
                               "sample": f"the first {r // P} of the {I} workload instances x {P} samples = {r} rollouts, one "
                                         f"full greedy episode ({t:.1f} s); {desc}"}
    print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
