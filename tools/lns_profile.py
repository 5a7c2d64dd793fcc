#!/usr/bin/env python
"""Where an LNS iteration's time goes (Solomon C103, 20 sub-problems, K1 overrides), for the all-device loop
(LNS.search_device: selection / destruction / import on tour buffers) and for the host-list loop (LNS.search: export_sol
-> Python destructor -> import_sol).  Phases are bracketed by device synchronisation, so each figure is host + device
time of that phase.

    python tools/lns_profile.py [iterations] [sub-problems per iteration = 20]"""
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import jampr_plus_b200.lns as L  # noqa: E402
from jampr_plus_b200 import Policy, RPEnv, driver  # noqa: E402
from jampr_plus_b200.lns import LNS, read_solomon  # noqa: E402
from jampr_plus_b200.weights import random_state_dict  # noqa: E402

warnings.simplefilter("ignore")
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_sub = int(sys.argv[2]) if len(sys.argv) > 2 else 20
inst = read_solomon(os.path.join(ROOT, "tests", "data", "c103.txt"))
env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=n_sub, inference=True,
            tour_graph_update_step=2, device="cuda")
pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
pol.load_state_dict(random_state_dict(0))
pol.eval()
T = {}


def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize()
        t = time.time()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        T[name] = T.get(name, 0) + time.time() - t
        return r
    return w


def report(title, tot):
    print(f"{title}: {tot:.3f} s for {n_it} iterations of {n_sub} sub-problems ({1e3 * tot / n_it:.2f} ms each, "
          f"{n_it * n_sub / tot:.0f} repairs/s)")
    for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
        print(f"  {k:14s} {1e3 * v / n_it:7.3f} ms  {100 * v / tot:5.1f} %")
    print(f"  {'(untimed)':14s} {1e3 * (tot - sum(T.values())) / n_it:7.3f} ms")


org = dict(destruct=env.destruct, import_sol=env.import_sol, export_sol=env.export_sol, true_cost=env.true_cost,
           gather_best=env.gather_best, rollout=driver.rollout)
# ---- all-device loop
LNS(inst, pol, env, seed=1234).search_device(num_steps=60)            # warm-up: library, graphs of the repair start steps
env.destruct = timed("destruct", org["destruct"])
env.true_cost = timed("true_cost", org["true_cost"])
env.gather_best = timed("gather_best", org["gather_best"])
L.rollout = timed("rollout", org["rollout"])
lns = LNS(inst, pol, env, seed=1234)
t0 = time.time()
lns.search_device(num_steps=n_it)
report("all-device loop (RPEnv.destruct)", time.time() - t0)
# ---- host-list loop
T.clear()
env.destruct, env.true_cost, env.gather_best = org["destruct"], org["true_cost"], org["gather_best"]
LNS(inst, pol, env, seed=1234).search(num_steps=3)
env.import_sol = timed("import_sol", org["import_sol"])
env.export_sol = timed("export_sol", org["export_sol"])
env.true_cost = timed("true_cost", org["true_cost"])
lns2 = LNS(inst, pol, env, seed=1234)
lns2.ds.destruct = timed("destructor", lns2.ds.destruct)
t0 = time.time()
lns2.search(num_steps=n_it)
report("host-list loop (export_sol -> Destructor -> import_sol)", time.time() - t0)
