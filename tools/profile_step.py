#!/usr/bin/env python
"""Full-size eager policy / environment steps for profiling (BASELINE config 2: 4096 workload instances x 16 samples):

    ncu --set full --clock-control none --import-source on -k regex:decoder_tc_kernel -s 20 -c 1 -o out python tools/profile_step.py 30

Launches are eager (no CUDA graph) so that ncu's -k / -s / -c select them one by one."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from jampr_plus_b200 import Policy, RPEnv, _cabi  # noqa: E402
from jampr_plus_b200.weights import random_state_dict  # noqa: E402

warnings.simplefilter("ignore")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
z = np.load(os.path.join(ROOT, "tests", "golden", "workload_all_1_low_4096.npz"))
arr = {k: torch.from_numpy(np.ascontiguousarray(z[k])).float().cuda()
       for k in ("coords", "demands", "tw", "service_time", "org_service_horizon")}
env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=16, inference=True, device="cuda")
env.seed(1235)
pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16", tc_variant=variant)
pol.load_state_dict(random_state_dict(0))
env.load_arrays(arr, int(z["max_vehicle_number"][0]), check=False)
env.reset()
st = env.state
st.ensure_policy_buffers(False, split=True)
w = pol.encode_static(env)
lib, d, p, S = pol._lib, env._dims, env._inst, env._stream()
step = env._step
for _ in range(steps):
    _cabi.check(lib.jampr_policy_step(d, p, w, st.struct(), 1, step, 0, 0, None, S), "jampr_policy_step")
    _cabi.check(lib.jampr_env_step(d, p, st.struct(), None, step, S), "jampr_env_step")
    step += 1
torch.cuda.synchronize()
print("done", steps, "steps; mean total so far", float(st["total"].mean()))
