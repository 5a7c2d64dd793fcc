"""Per-phase time line of the fused policy-step kernel at full load (65,536 rollouts), from clock64() stamps taken by
thread 0 of every 64th CTA in the last live step of an episode.  Needs the instrumented library:
    make -C jampr_plus_b200/csrc clean all EXTRA=-DJAMPR_TIMELINE
Prints the mean time a CTA spends between consecutive stamps (microseconds at the SM clock) — wall time of a CTA
that shares its SM with a second one, not issue time."""
import ctypes
import json
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, ".")
NAMES = {1: "prologue (state, tables, links)", 2: "h0 load -> tiles / TMEM skip"}
for l in range(6):
    kind = "kNN" if l in (2, 5) else "tour"
    NAMES[3 + 3 * l] = f"layer {l} ({kind}): root MMA issue + aggregation"
    NAMES[4 + 3 * l] = f"layer {l}: wait lin_rel weights + MMA"
    NAMES[5 + 3 * l] = f"layer {l}: epilogue (GELU, skip, LayerNorm)"
NAMES.update({21: "projection MMAs + per-tour means", 22: "node-embedding epilogue (+ static_emb)", 23: "dec: tour mat-vec, column stats, tour_emb",
              24: "dec: ctxt mat-vec", 25: "dec: glimpse-key mat-vec", 26: "dec: glimpse scores (per node)",
              27: "dec: softmax, glimpse weights", 28: "dec: weighted sums (gbar)", 29: "dec: glimpse-value mat-vec",
              30: "dec: logit-key mat-vec", 31: "dec: logits (per node)", 32: "dec: log-softmax + selection"})


def main(n_inst=4096, mhz=1965.0):
    from jampr_plus_b200 import Policy, RPEnv, _cabi
    from jampr_plus_b200.driver import rollout
    from jampr_plus_b200.synth import instances_to_arrays, sample_instances
    from jampr_plus_b200.weights import random_state_dict
    lib = _cabi.load_library()
    if not hasattr(lib, "jampr_debug_timeline"):
        raise SystemExit("library was not built with -DJAMPR_TIMELINE")
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                             "workload_all_1_low_4096.npz"))
    arr = {k: z[k][:n_inst] for k in z.files if z[k].ndim > 0}
    dev = {k: torch.from_numpy(np.ascontiguousarray(arr[k])).to(torch.float32).cuda()
           for k in ("coords", "demands", "tw", "service_time", "org_service_horizon")}
    env = RPEnv(max_concurrent_vehicles=3, k_nbh_frac=0.25, pomo=True, num_samples=16, inference=True, device="cuda")
    env.seed(1235)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision="fp16",
                 tc_variant=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    pol.load_state_dict(random_state_dict(0))
    pol.eval()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(2):
            env.load_arrays(dev, int(arr["max_vehicle_number"][0]), check=False)
            env.reset()
            rollout(pol, env, check=False)
    out = np.zeros((1024, 40), dtype=np.uint32)
    rc = lib.jampr_debug_timeline(out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    rows = out[(out[:, 32] != 0) & (out[:, 0] != 0)].astype(np.int64)
    d = (rows[:, 1:33] - rows[:, 0:32]) & 0xFFFFFFFF
    total = ((rows[:, 32] - rows[:, 0]) & 0xFFFFFFFF).mean() / mhz
    print(f"# {len(rows)} sampled CTAs, mean CTA lifetime {total:.1f} us (two CTAs per SM)")
    res = {}
    for k in range(1, 33):
        us = d[:, k - 1].mean() / mhz
        res[NAMES[k]] = us
        print(f"{us:7.2f} us {100 * us / total:5.1f} %  {NAMES[k]}")
    grp = {"aggregation (6 layers)": sum(v for n, v in res.items() if "aggregation" in n),
           "MMA / weight waits (6 layers)": sum(v for n, v in res.items() if "wait lin_rel" in n),
           "layer epilogues": sum(v for n, v in res.items() if ": epilogue" in n),
           "decoder": sum(v for n, v in res.items() if n.startswith("dec:"))}
    print(json.dumps({k: round(v, 2) for k, v in grp.items()}))


if __name__ == "__main__":
    main()
