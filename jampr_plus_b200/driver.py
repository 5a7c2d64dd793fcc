"""Whole construction rollouts without host round trips.

``rollout()`` replaces the reference's Python loop ``while not done: action = policy(obs); obs = env.step(action)``
(``lib/model/training.py:72-77``, ``lib/lns/lns.py:258-261``) by one ``jampr_rollout`` call that enqueues every
step on the current stream; termination lives in a device control word.  ``eval_episode`` mirrors
``training.py:54-96`` (best-of-samples reduction included)."""
from typing import Dict, List, Optional, Tuple

import torch

from . import _cabi
from .routing.formats import RPInstance

__all__ = ["rollout", "rollout_sharded", "eval_episode"]


def _enqueue(policy, env, w, n):
    st = env.state
    rc = policy._lib.jampr_rollout(env._dims, env._inst, w, st.struct(), int(env._step), int(env._graph_fresh),
                                   policy._decode_code(), int(policy.sampling_seed), int(n), env._stream())
    return _cabi.check(rc, "jampr_rollout")


def _graph_key(policy, env, n):
    """Everything the captured kernels take BY VALUE: every device pointer of the instance tables, of the state and of
    both weight blobs, the dims struct, and the launch arguments.  No ``id()``: an address that is reused after a
    reallocation only matches when every other baked-in address and every dimension matches too, in which case the
    captured launches are the right ones.  The graphs live on the DeviceState object (dropped with its buffers)."""
    st = env.state
    inst = tuple(0 if getattr(env._inst, f) is None else int(getattr(env._inst, f)) for f, _ in env._inst._fields_)
    dims = tuple(int(getattr(env._dims, f)) for f, _ in env._dims._fields_)
    state = tuple(sorted((name, t.data_ptr()) for name, t in st.t.items()))
    return (inst, dims, state, policy._blob.data_ptr(), 0 if policy._blob_bf16 is None else policy._blob_bf16.data_ptr(),
            policy._blob_version, policy.precision, int(policy.tc_variant), int(env._step), int(env._graph_fresh), policy._decode_code(),
            int(policy.sampling_seed), int(n), int(st["ctl"].data_ptr()))


_MAX_GRAPHS_PER_STATE = 64      # e.g. one graph per start step of the LNS repair rollouts; least recently used goes first


def _replay_or_capture(policy, env, w, n):
    st = env.state
    cache = st.graphs
    key = _graph_key(policy, env, n)
    hit = cache.pop(key, None)
    if hit is None:
        rc = _enqueue(policy, env, w, n)            # eager first run (also sets the kernels' smem attributes)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):                  # capture launches nothing: the eager run above is the result
            rc_cap = _enqueue(policy, env, w, n)
        assert rc_cap == rc
        hit = (g, rc)
        while len(cache) >= _MAX_GRAPHS_PER_STATE:
            cache.pop(next(iter(cache)))           # dicts keep insertion order: the oldest entry
    else:
        hit[0].replay()
    cache[key] = hit                               # (re)insert as the most recent entry
    return hit[1]


def rollout(policy, env, max_steps: Optional[int] = None, check: bool = True,
            graph: bool = False) -> Tuple[torch.Tensor, Dict]:
    """Run the episode currently held by ``env`` (after reset() / import_sol()) to termination on the device.

    ``graph=True`` captures the whole multi-step loop (3 kernels per step, ~130 steps) into a CUDA graph the first
    time a (device buffers, dims, weights, start step, decode type) combination is seen and replays it afterwards:
    one host call per episode, which is what matters for the small batches of the LNS repair loop (lns.py:236-275).
    Requires stable buffers (RPEnv reuses its device buffers for equally shaped batches), greedy decoding or a
    fixed sampling seed.

    Returns (total cost (bs,), info) with the reference's info keys."""
    assert env._is_reset, "call env.reset() or env.import_sol() first"
    st = env.state
    st.ensure_policy_buffers(policy.return_logp, split=policy.precision != "fp32")
    w = policy.encode_static(env)
    cap = env.graph_size + env.max_vehicle_number + 1
    n = cap - env._step + 1 if max_steps is None else max_steps
    rc = _replay_or_capture(policy, env, w, n) if graph else _enqueue(policy, env, w, n)
    info: Dict = {}
    if check:
        done_step = int(st["ctl"][_cabi.CTL_DONE_STEP].item())      # the one device->host sync of the rollout
        env._raise_status()
        if done_step >= 0:
            env._step = done_step + 1
            env._has_instance = False
            env._is_reset = False
            env._done = True
        else:
            env._step += rc
        info = {"done_step": done_step, "steps_enqueued": rc}
    return st["total"], info


def rollout_sharded(policy, env, group=None, reduce_max=None, graph: bool = False) -> Tuple[torch.Tensor, Dict]:
    """``rollout`` for one shard of a batch that is split over several ranks (jampr_plus_b200/sharding.py), with the
    reference's termination semantics for the UNION batch: the reference stops when every rollout of the whole batch
    is complete (env.py:281) and the reported ``total_cost`` depends on that step (the return legs of vehicles still
    en route are double counted).  Protocol (include/jampr_b200.h, JAMPR_CTL_HOLD): the shard runs
    until its own finishing condition and parks; one 4-byte all_reduce(MAX) yields the union's finishing step; the
    parked step is resumed and the remaining steps are enqueued, so the terminal accounting happens at the same
    step on every rank.  Tours never depend on this; ``RPEnv.true_cost()`` does not either.

    ``reduce_max``: callable int -> int replacing the collective (tests emulate several ranks in one process).
    ``graph=True``: the first phase (every step up to the shard's own finishing step: ~130 steps x 3 kernels) is
    replayed as one captured CUDA graph like in ``rollout``; the short second phase (resume + the steps up to the
    union's finishing step, a data-dependent count that is usually 0..2) is enqueued eagerly."""
    from .sharding import global_done_step
    assert env._is_reset, "call env.reset() or env.import_sol() first"
    st = env.state
    st.ensure_policy_buffers(policy.return_logp, split=policy.precision != "fp32")
    w = policy.encode_static(env)
    ctl = st["ctl"]
    ctl[_cabi.CTL_HOLD] = 1
    cap = env.graph_size + env.max_vehicle_number + 1
    n1 = cap - env._step + 1
    rc = _replay_or_capture(policy, env, w, n1) if graph else _enqueue(policy, env, w, n1)
    local = int(ctl[_cabi.CTL_LOCAL_DONE].item())
    assert local >= 0, "shard did not reach a finishing condition"
    final = int(reduce_max(local)) if reduce_max is not None else global_done_step(local, env.device, group)
    ctl[_cabi.CTL_FINAL_STEP] = final
    _cabi.check(policy._lib.jampr_env_resume(env._dims, env._inst, st.struct(), int(local), env._stream()),
                "jampr_env_resume")
    if final > local:
        env._step = local + 1
        env._graph_fresh = (local <= env.max_concurrent_vehicles + 1) or (local % env.tour_graph_update_step == 0)
        _enqueue(policy, env, w, final - local)
    done_step = int(ctl[_cabi.CTL_DONE_STEP].item())
    assert done_step == final, (done_step, final)
    env._raise_status()
    env._step = done_step + 1
    env._has_instance = False
    env._is_reset = False
    env._done = True
    return st["total"], {"done_step": done_step, "local_done_step": local, "steps_enqueued": rc + max(final - local, 0)}


def eval_episode(data: List[RPInstance], env, policy, sampling: bool = False, **kwargs):
    """One evaluation episode (training.py:54-96): returns (cost per instance, info)."""
    policy.eval()
    policy.set_decode_type("sampling" if sampling and not env.pomo else "greedy")
    policy.reset_static()
    env.load_data(data)
    env.reset()
    total, info = rollout(policy, env)
    info["current_total_cost"] = total.cpu().numpy()
    info["k_used"] = env.k_used.cpu().numpy()
    costs = total.cpu()
    if env.num_samples > 1:
        best_cost, best_idx, _ = env.gather_best()
        bi = best_idx.long().cpu()
        rng = torch.arange(bi.numel())
        info = {k: (v.reshape(-1, env.num_samples)[rng.numpy(), bi.numpy()] if hasattr(v, "reshape") else v)
                for k, v in info.items()}
        costs = best_cost.cpu()
    return costs, info
