"""Per-phase time line of decoder_tc_kernel at full load (65,536 rollouts, mid-episode step), from clock64() stamps
taken by thread 0 of every CTA.  Needs the instrumented library:
    make -C jampr_plus_b200/csrc clean all EXTRA=-DJAMPR_TIMELINE
Usage: python tools/dtimeline.py [steps=30]"""
import ctypes
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from jampr_plus_b200 import Policy, RPEnv, _cabi  # noqa: E402
from jampr_plus_b200.weights import random_state_dict  # noqa: E402

NAMES = ["stage T (tour-embedding GEMMs, drains, query hand-over)", "(block barrier)", "stage A (query + ctxt GEMM)",
         "stage B (glimpse-key GEMMs + drains)", "phase C (glimpse per rollout)", "stage D (glimpse-value GEMMs)",
         "stage E (logit-key GEMM + drain)", "phase F (logits + selection, warp 0 only)"]


def main(steps=30, mhz=1965.0):
    warnings.simplefilter("ignore")
    lib = _cabi.load_library()
    if not hasattr(lib, "jampr_debug_dtimeline"):
        raise SystemExit("library was not built with -DJAMPR_TIMELINE")
    z = np.load(os.path.join(ROOT, "tests", "golden", "workload_all_1_low_4096.npz"))
    arr = {k: torch.from_numpy(np.ascontiguousarray(z[k])).float().cuda()
           for k in ("coords", "demands", "tw", "service_time", "org_service_horizon")}
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=16, inference=True, device="cuda")
    env.seed(1235)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16")
    pol.load_state_dict(random_state_dict(0))
    env.load_arrays(arr, int(z["max_vehicle_number"][0]), check=False)
    env.reset()
    st = env.state
    st.ensure_policy_buffers(False, split=True)
    w = pol.encode_static(env)
    d, p, S = env._dims, env._inst, env._stream()
    step = env._step
    for _ in range(steps):
        _cabi.check(lib.jampr_policy_step(d, p, w, st.struct(), 1, step, 0, 0, None, S), "jampr_policy_step")
        _cabi.check(lib.jampr_env_step(d, p, st.struct(), None, step, S), "jampr_env_step")
        step += 1
    torch.cuda.synchronize()
    out = np.zeros((1024, 16), dtype=np.uint32)
    assert lib.jampr_debug_dtimeline(out.ctypes.data_as(ctypes.c_void_p)) == 0
    rows = out[(out[:, 8] != 0) & (out[:, 0] != 0)].astype(np.int64)
    dd = (rows[:, 1:9] - rows[:, 0:8]) & 0xFFFFFFFF
    total = ((rows[:, 8] - rows[:, 0]) & 0xFFFFFFFF).mean() / mhz
    mask = (st["mask"].view(-1, 78) == 0).sum(1).float()
    print(f"# {len(rows)} CTAs, mean CTA lifetime {total:.1f} us; feasible actions per rollout: mean {float(mask.mean()):.1f} "
          f"max {int(mask.max())}")
    for k in range(8):
        us = dd[:, k].mean() / mhz
        print(f"{us:8.2f} us {100 * us / total:5.1f} %  {NAMES[k]}")
    # inside stage T (thread 0; stamps 9..15 relative to the start of the kernel body, stamp 0)
    fine = ["L2 prefetches, barrier init, TMEM allocation", "first A tile (HBM load + tile write)", "item 0 issued",
            "items 1, 2 issued, slot 0 complete", "slot 0 drained", "slots 1 .. K-1", "query hand-over"]
    prev = rows[:, 0]
    for k, name in zip(range(9, 16), fine):
        us = (((rows[:, k] - prev) & 0xFFFFFFFF).mean()) / mhz
        print(f"    {us:7.2f} us  stage T: {name}")
        prev = rows[:, k]


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 30)
