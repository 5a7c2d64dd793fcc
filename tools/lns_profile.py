import os, sys, time, warnings
sys.path.insert(0, os.getcwd())
import torch
from jampr_plus_b200 import Policy, RPEnv
from jampr_plus_b200.lns import LNS, read_solomon
from jampr_plus_b200.weights import random_state_dict
from jampr_plus_b200 import driver
warnings.simplefilter("ignore")
inst = read_solomon("tests/data/c103.txt")
env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=20, inference=True, tour_graph_update_step=2, device="cuda")
pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16"); pol.load_state_dict(random_state_dict(0)); pol.eval()
lns = LNS(inst, pol, env, seed=1234)
lns.search(num_steps=3)
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.time(); r = fn(*a, **k); torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.time() - t; return r
    return w
env.import_sol = timed("import_sol", env.import_sol)
env.export_sol = timed("export_sol", env.export_sol)
env.true_cost = timed("true_cost", env.true_cost)
lns.ds.destruct = timed("destruct", lns.ds.destruct)
import jampr_plus_b200.lns as L
L.rollout = timed("rollout", driver.rollout)
lns2 = LNS(inst, pol, env, seed=1234); lns2.ds.destruct = timed("destruct", lns2.ds.destruct)
t0 = time.time(); lns2.search(num_steps=200); tot = time.time() - t0
print("total %.3f s for 200 iterations (%.2f ms each)" % (tot, tot * 5))
for k, v in sorted(T.items(), key=lambda kv: -kv[1]): print("%-12s %.3f s  %.1f %%" % (k, v, 100 * v / tot))
