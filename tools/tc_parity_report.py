"""Teacher-forced comparison of the tensor-core policy step (bf16 / fp16 operands) with the golden traces of the
reference: log-probability errors, greedy-decision agreement by reference top-2 gap, and the embedding errors
against the fp32 CUDA path.  Prints one JSON line per (case, precision).  Needs a GPU."""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.replay import CASES, env_args, load_case, start_nodes_of  # noqa: E402


def make(g, precision):
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.weights import random_state_dict
    env = RPEnv(inference=True, device="cuda", **env_args(g["env_kwargs"]))
    env.seed(1)
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cuda", precision=precision, allow_reduced_accuracy=True)
    pol.load_state_dict(random_state_dict(int(g["policy_seed"])))
    pol.eval()
    pol.return_logp = True
    env.load_data(g["instances"], dist_mat=torch.from_numpy(g["static_dist"]))
    return env, pol


def run(name, precision, max_steps=None):
    g = load_case(name)
    env, pol = make(g, precision)
    envr, polr = make(g, "fp32")
    sn = start_nodes_of(g)
    sn = None if sn is None else sn.to(torch.int32)
    obs = env.reset(start_nodes=sn)
    obr = envr.reset(start_nodes=sn)
    T = g["trace_action"].shape[0] if max_steps is None else min(max_steps, g["trace_action"].shape[0])
    errs, rels, gaps, agree, herr, neerr, ferr = [], [], [], [], [], [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for t in range(T):
            pol(obs)
            polr(obr)
            lp = pol.log_probs(env).cpu().numpy()
            lr = polr.log_probs(envr).cpu().numpy()
            ref = g["trace_logp"][t]
            fin = np.isfinite(ref)
            assert np.array_equal(np.isfinite(lp), fin), (name, precision, t, "support")
            errs.append(np.abs(lp - ref)[fin].max())
            rels.append((np.abs(lp - ref)[fin] / np.maximum(1.0, np.abs(ref[fin]))).max())
            ferr.append(np.abs(lr - ref)[fin].max())
            srt = np.sort(ref, -1)
            gaps.append(srt[:, -1] - srt[:, -2])
            agree.append((env.state["action"].cpu().numpy() == g["trace_action"][t]).all(-1)
                         if g["decode"] == "greedy" else np.ones(lp.shape[0], bool))
            herr.append(float((env.state["h_fin"] - envr.state["h_fin"]).abs().max()))
            neerr.append(float((env.state["ne"] - envr.state["ne"]).abs().max()))
            act = torch.from_numpy(g["trace_action"][t].astype(np.int64))
            obs, _, done, _ = env.step(act)
            obr, _, _, _ = envr.step(act)
            if done:
                break
    errs, gaps, agree = np.array(errs), np.concatenate(gaps), np.concatenate(agree)
    out = {"case": name, "precision": precision, "steps": len(errs), "decisions": int(agree.size),
           "logp_abs_err_max": float(errs.max()), "logp_rel_err_max(|ref|>=1 floor)": float(np.max(rels)), "logp_abs_err_median_of_step_max": float(np.median(errs)),
           "fp32_path_logp_abs_err_max": float(np.max(ferr)),
           "h_fin_abs_err_max": float(np.max(herr)), "ne_abs_err_max": float(np.max(neerr)),
           "mismatch_total": int((~agree).sum())}
    for thr in (1e-3, 3e-3, 1e-2, 3e-2, 1e-1):
        sel = gaps > thr
        out[f"mismatch_gap>{thr:g}"] = int((~agree[sel]).sum())
        out[f"decisions_gap>{thr:g}"] = int(sel.sum())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16", "fp16"]
    cases = sys.argv[2].split(",") if len(sys.argv) > 2 else CASES
    for c in cases:
        for pr in precs:
            run(c, pr)
