#!/usr/bin/env python
"""Counterpart of the reference's ``run_lns.py`` (hydra entry point, run_lns.py:9-19 -> lib/lns/lns.py) on the B200 path:

    python run_lns.py data_path=tests/data/c103.txt time_limit=60 [config=my.yaml] [key=value ...]

loads the Solomon / Gehring-Homberger instance, picks the checkpoint by instance type and time-window fraction
(type_twf_map_model; seeded stand-in weights when the file is absent, as in this image), and runs
construct -> [select -> destroy -> repair]* for ``time_limit`` seconds or ``num_steps`` iterations.  Prints one JSON
line: incumbent cost (environment total and batch-independent tour cost, original units), iterations, tours.
Several processes under torchrun shard the sub-problems over the GPUs (BASELINE config 5)."""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main(argv):
    import torch
    import torch.distributed as dist
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.checkpoint import load_checkpoint
    from jampr_plus_b200.config import load_config, select_checkpoint
    from jampr_plus_b200.lns import LNS, Destructor, read_solomon
    from jampr_plus_b200.weights import random_state_dict
    path = None
    ov = []
    for a in argv:
        if a.startswith("config="):
            path = a.split("=", 1)[1]
        else:
            ov.append(a)
    cfg = load_config(path, ov)
    if cfg.no_warnings:
        warnings.filterwarnings("ignore")
    if not cfg.data_path:
        raise SystemExit("data_path=<instance file> is required")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    inst = read_solomon(cfg.data_path)
    ec = cfg.env_cfg
    env = RPEnv(max_concurrent_vehicles=ec.max_concurrent_vehicles, k_nbh_frac=ec.k_nbh_frac, pomo=cfg.pomo_inference,
                pomo_single_start_node=cfg.pomo_single_start_node, num_samples=cfg.num_samples, inference=True,
                tour_graph_update_step=ec.tour_graph_update_step, device=dev, debug=cfg.debug_lvl)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device=dev, precision=cfg.precision)
    ckpt = select_checkpoint(cfg, inst.type, inst.tw_frac)
    weights = "checkpoint " + ckpt
    if os.path.exists(ckpt):
        pol.load_state_dict(load_checkpoint(ckpt)[0])
    else:
        weights = f"seeded stand-in weights ({ckpt} not found)"
        pol.load_state_dict(random_state_dict(0))
    pol.eval()
    dc = cfg.ds_cfg
    assert dc.destruction_mode == "random_from_route"
    ds = Destructor(num_partial=ec.max_concurrent_vehicles, rm_partial_mode=dc.rm_partial_mode,
                    rm_complete_mode=dc.rm_complete_mode, frac_rm_nodes_from_partial=dc.frac_rm_nodes_from_partial,
                    frac_rm_complete_routes=dc.frac_rm_complete_routes, seed=cfg.seed)
    device_loop = bool(cfg.device_loop) and (cfg.selection_mode == "random" or world > 1)
    lns = LNS(inst, pol, env, num_best=cfg.num_best, num_other=cfg.num_other, add_best=cfg.add_best, destructor=ds,
              seed=cfg.seed, selection_mode=cfg.selection_mode, ls_step=int(cfg.ls_step))
    kw = dict(time_limit=float(cfg.time_limit)) if cfg.time_limit else dict(num_steps=int(cfg.num_steps))
    t0 = time.time()
    if device_loop:
        plan, cost, true_cost = lns.search_device(**kw)
        tours = [[int(x) for x in r[r > 0]] for r in plan.cpu().numpy() if (r > 0).any()]
    else:
        sol, cost, true_cost = lns.search(**kw)
        tours = [t for t in sol if len(t) > 1]
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"ID": cfg.ID, "instance": os.path.basename(cfg.data_path), "weights": weights,
                          "loop": "device" if device_loop else "host lists", "n_gpus": world, "seconds": time.time() - t0,
                          "iterations": lns.step, "incumbent_env_cost": cost, "incumbent_tour_cost": true_cost,
                          "n_tours": len(tours), "tours": tours}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1:])
