"""Batched CVRP-TW construction environment on the B200: host-side mirror of the reference ``RPEnv``.

Same constructor, ``load_data / reset / step / import_sol / export_sol`` surface and attribute names as
reference ``lib/routing/env.py:19-1297``; every state transition, candidate set and feasibility mask is
computed by libjampr_b200.so (include/jampr_b200.h) on ``device``.  There is no CPU implementation here:
a non-CUDA device raises.

Differences a caller can observe (all documented in INTEGRATION.md):
  * static data lives once per *instance* on the device (the reference replicates it per sample,
    env.py:1189-1202); ``coords``/``demands``/``tw`` therefore have leading dimension ``n_inst``.
  * the observation tuple keeps the reference fields, but the edge lists (``node_nbh_edges`` etc.) are only
    materialised when ``RPEnv(..., debug>=1)``; the policy reads the device state through ``obs.state``.
  * data-dependent errors are raised from per-rollout status words after the step (same exception types
    and messages as the reference: env.py:230-236, :254-258, :1020-1021).
"""
import math
import warnings
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .. import _cabi
from .formats import RPInstance, RPObs

__all__ = ["RPEnv", "DeviceState"]

# Debug aid (compute-sanitizer is not available on every pool): with GUARD_BYTES > 0 every device buffer the kernels
# write is carved out of a larger allocation with GUARD_BYTES of 0xA5 on both sides; ``check_guards`` verifies the
# pattern.  Set through JAMPR_DEBUG_GUARD=<bytes> or by the tests (tests/test_gpu_guards.py).
import os as _os

GUARD_BYTES = int(_os.environ.get("JAMPR_DEBUG_GUARD", "0"))


class _GuardedAlloc:
    """Allocator of zero-initialised device tensors, optionally fenced by guard bytes."""

    def __init__(self):
        self.fenced = []      # (label, raw uint8 tensor, payload bytes, guard bytes)

    def zeros(self, shape, dtype, device, label=""):
        if GUARD_BYTES <= 0:
            return torch.zeros(shape, dtype=dtype, device=device)
        g = (GUARD_BYTES + 255) // 256 * 256           # keeps the payload 256-byte aligned
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        raw = torch.full((n + 2 * g,), 0xA5, dtype=torch.uint8, device=device)
        self.fenced.append((label, raw, n, g))
        view = raw[g:g + n]
        view.zero_()
        return view.view(dtype).view(shape)

    def check(self):
        bad = []
        for label, raw, n, g in self.fenced:
            if not bool((raw[:g] == 0xA5).all()) or not bool((raw[g + n:] == 0xA5).all()):
                bad.append(label)
        if bad:
            raise RuntimeError(f"guard bytes overwritten around: {bad}")
        return len(self.fenced)


_STATE_SPEC = {
    # name: (dtype, shape as a lambda of dims)
    "visited": (torch.uint8, lambda s: (s.bs, s.N)),
    "pred": (torch.int16, lambda s: (s.bs, s.N)),
    "succ": (torch.int16, lambda s: (s.bs, s.N)),
    "plan": (torch.int16, lambda s: (s.bs, s.K_max, s.L)),
    "vflags": (torch.uint8, lambda s: (s.bs, s.K_max)),
    "row": (torch.int32, lambda s: (s.bs, s.K)),
    "nxt": (torch.int32, lambda s: (s.bs, s.K)),
    "cur": (torch.int32, lambda s: (s.bs, s.K)),
    "row_last": (torch.int32, lambda s: (s.bs, s.K)),
    "cap": (torch.float32, lambda s: (s.bs, s.K)),
    "time": (torch.float32, lambda s: (s.bs, s.K)),
    "cttd": (torch.float32, lambda s: (s.bs, s.K)),
    "total": (torch.float32, lambda s: (s.bs,)),
    "cost": (torch.float32, lambda s: (s.bs,)),
    "status": (torch.int32, lambda s: (s.bs,)),
    "allvis": (torch.uint8, lambda s: (s.bs,)),
    "nvis": (torch.int32, lambda s: (s.bs,)),
    "nbh": (torch.int16, lambda s: (s.bs, s.K, s.k)),
    "mask": (torch.uint8, lambda s: (s.bs, s.K, s.k)),
    "action": (torch.int32, lambda s: (s.bs, 2)),
    "ll": (torch.float32, lambda s: (s.bs,)),
    "entropy": (torch.float32, lambda s: (s.bs,)),
    "ctl": (torch.int32, lambda s: (_cabi.CTL_WORDS,)),
}
# policy-side buffers, allocated on first use by the policy (they dominate memory: ~1.5 KB * N per rollout)
_POLICY_SPEC = {
    "logp": (torch.float32, lambda s: (s.bs, s.K * s.k)),
    "h_fin": (torch.float32, lambda s: (s.bs, s.N, 128)),
    "ne": (torch.float32, lambda s: (s.bs, s.N, 256)),
    "ne_stats": (torch.float32, lambda s: (s.bs, 2, 256)),
}
# split decoder of the tensor-core path (csrc/decoder_tc.cu): hand-over buffers between the per-rollout GNN kernel and
# the 128-rollouts-per-CTA decoder kernel; allocated when a 16-bit policy drives the env (up to four concurrent vehicles)
_SPLIT_SPEC = {
    "ne16": (torch.int16, lambda s: (s.bs, s.N, 256)),
    "dq": (torch.float32, lambda s: (s.bs, 512)),
    "dte": (torch.float32, lambda s: (s.bs, 4, 256)),
    "dqp": (torch.float32, lambda s: (s.bs, 4, 256)),
    "dgbar": (torch.float32, lambda s: (s.bs, 4, 256)),
    "dlp": (torch.float32, lambda s: (s.bs, 256)),
}
# large-graph path only (N > 128: the graph spans several 128-row tiles, csrc/gnn_large.cu)
_LARGE_SPEC = {
    "hbuf": (torch.int16, lambda s: (2, s.bs, (s.N + 127) // 128 * 128, 128)),
    "hskip": (torch.float32, lambda s: (s.bs, (s.N + 127) // 128 * 128, 128)),
    "stats_part": (torch.float32, lambda s: (s.bs, (s.N + 127) // 128, 2, 256)),
}


class DeviceState:
    """Per-rollout episode state in HBM (struct-of-arrays, layout of jampr_state_t)."""

    def __init__(self, bs: int, N: int, K: int, k: int, K_max: int, L: int, device: torch.device):
        self.bs, self.N, self.K, self.k, self.K_max, self.L = bs, N, K, k, K_max, L
        self.device = device
        self.t = {}
        self._alloc = _GuardedAlloc()
        for name, (dt, shp) in _STATE_SPEC.items():
            self.t[name] = self._alloc.zeros(shp(self), dt, device, name)
        self.want_logp = False
        self._struct = None
        self.graphs = {}      # captured CUDA graphs of rollouts over these buffers (driver._replay_or_capture)

    def ensure_policy_buffers(self, want_logp: bool = False, split: bool = False):
        changed = False
        if split and self.K <= 4:      # (graphs of several tiles: the layer-wise GNN feeds the same decoder)
            for name, (dt, shp) in _SPLIT_SPEC.items():
                if name not in self.t:
                    self.t[name] = self._alloc.zeros(shp(self), dt, self.device, name)
                    changed = True
        for name, (dt, shp) in _POLICY_SPEC.items():
            if name == "logp" and not (want_logp or self.want_logp):
                continue
            if name not in self.t:
                self.t[name] = self._alloc.zeros(shp(self), dt, self.device, name)
                changed = True
        if self.N > 128:
            for name, (dt, shp) in _LARGE_SPEC.items():
                if name not in self.t:
                    self.t[name] = self._alloc.zeros(shp(self), dt, self.device, name)
                    changed = True
        self.want_logp = self.want_logp or want_logp
        if changed:
            self._struct = None

    def struct(self) -> "_cabi.State":
        if self._struct is None:
            s = _cabi.State()
            for name in _cabi.STATE_FIELDS:
                setattr(s, name, _cabi.ptr(self.t.get(name)))
            self._struct = s
        return self._struct

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.t[name]


class RPEnv:
    """Drop-in for reference ``lib.routing.RPEnv`` (env.py:19-1297) backed by libjampr_b200.so."""
    OBSERVATION_SPACE = {
        "node_features": 7,
        "tour_features": 5,
    }

    def __init__(self,
                 check_feasibility: bool = False,
                 max_concurrent_vehicles: int = 3,
                 k_nbh_frac: float = 0.25,
                 pomo: bool = False,
                 pomo_nbh_frac: float = 0.25,
                 pomo_single_start_node: bool = False,
                 num_samples: int = 1,
                 enable_render: bool = False,
                 plot_save_dir: Optional[str] = None,
                 device: Union[torch.device, str] = "cuda",
                 fp_precision: torch.dtype = torch.float,
                 inference: bool = False,
                 tour_graph_update_step: int = 1,
                 debug: Union[bool, int] = False,
                 ):
        self.check_feasibility = check_feasibility
        self.max_concurrent_vehicles = max_concurrent_vehicles
        self.k_nbh_frac = k_nbh_frac
        self.pomo = pomo
        if self.pomo:
            if num_samples <= 1:
                warnings.warn("POMO should use num_samples > 1")
            self._num_samples = num_samples
        elif inference:
            self._num_samples = num_samples
        else:
            self._num_samples = 1
        self.pomo_nbh_frac = pomo_nbh_frac
        self.pomo_single_start_node = pomo_single_start_node
        if enable_render:
            raise NotImplementedError("rendering is outside the accelerated path (reference visualization.py)")
        self.enable_render = False
        self.plot_save_dir = plot_save_dir
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("jampr_plus_b200.RPEnv runs on a CUDA device only (no CPU fallback)")
        if fp_precision != torch.float:
            raise NotImplementedError("the CUDA environment computes in fp32 (reference default, env.py:41)")
        self.fp_precision = fp_precision
        assert tour_graph_update_step != 0
        if tour_graph_update_step < 0:
            # the reference only asserts != 0 (env.py:118); a negative period has no meaning in `_step % period == 0`
            # beyond "Python's sign convention", and the fused device loop rejects it: refuse it in one place
            raise ValueError("tour_graph_update_step must be a positive integer")
        self.tour_graph_update_step = tour_graph_update_step
        # the reference keeps the dense distance matrix only in inference mode (env.py:369-372); the kernels
        # always use it, which yields the same values (same formula, challenge_utils.py:35)
        self.inference = inference
        self.debug_lvl = 2 if isinstance(debug, bool) and debug else int(debug)
        if self.debug_lvl > 0:
            self.check_feasibility = True
        self._lib = _cabi.load_library()
        self._gen = torch.Generator(device=self.device)
        self._seed_value, self._n_resets = 0, 0
        self.clear_cache()

    # ------------------------------------------------------------------ plumbing
    def seed(self, seed: int) -> List[int]:
        """Reference seeds the global torch RNG (env.py:142-144); the start-node draw here uses a
        device generator owned by the env, seeded the same way."""
        torch.manual_seed(seed)
        self._gen.manual_seed(seed)
        self._seed_value, self._n_resets = int(seed), 0
        return [seed]

    def check_guards(self) -> int:
        """Debug aid: verify the guard bytes around every kernel-written buffer (GUARD_BYTES > 0); returns the
        number of fenced buffers, raises RuntimeError naming the buffers whose fences were overwritten."""
        n = 0
        for a in (getattr(self, "_table_alloc", None), getattr(getattr(self, "state", None), "_alloc", None)):
            if a is not None:
                n += a.check()
        return n

    def clear_cache(self):
        self.bs = None
        self.n_inst = None
        self.coords = self.demands = self.tw = self.service_time = None
        self.graph_size = None
        self.org_service_horizon = None
        self.max_vehicle_number = None
        self.vehicle_capacity = None
        self.service_horizon = None
        self.time_to_depot = None
        self._dist_mat = None
        self.k_nbh_size = None
        self.ordered_idx = None
        self._knn = self._knn_w = None
        self._static_emb = self._h0 = None
        self._tc_tab = self._tc_dtab = None
        self._lg_h16 = self._lg_skip = None
        self._inst_status = None
        self._dims = None
        self._inst = None
        self.state = None
        self._has_instance = False
        self._is_reset = False
        self._step = None
        self._graph_fresh = True
        self._static_version = getattr(self, '_static_version', 0) + 1
        self._table_key = None
        self._max_step_buf = None
        self._destroy_buf = None

    def _stream(self):
        return _cabi.stream_ptr(self.device)

    # ------------------------------------------------------------------ load_data (env.py:344-418)
    def load_data(self, batch: List[RPInstance], dist_mat: Optional[torch.Tensor] = None) -> None:
        """Load a list of RPInstances: uploads them and builds the per-instance tables on the device.
        ``dist_mat`` (n_inst, N, N), optional and not in the reference: a normalised transit matrix to use
        instead of evaluating the DIMACS formula on the device."""
        gs = int(batch[0].graph_size)
        assert np.all(np.array([x.graph_size for x in batch]) == gs)
        kv = batch[0].max_vehicle_number
        assert np.all(np.array([x.max_vehicle_number for x in batch]) == kv)
        assert np.all(np.array([x.vehicle_capacity for x in batch]) == 1)
        assert np.all(np.array([x.service_horizon for x in batch]) == 1)
        assert np.all(np.array([x.depot_idx[0] for x in batch]) == 0)
        self.vehicle_capacity = batch[0].vehicle_capacity
        self.service_horizon = batch[0].service_horizon
        stack = lambda key: torch.from_numpy(np.stack([np.asarray(x[key]) for x in batch])).to(torch.float32)
        self.load_arrays({"coords": stack("coords"), "demands": stack("demands"), "tw": stack("tw"),
                          "service_time": stack("service_time"),
                          "org_service_horizon": stack("org_service_horizon")}, int(kv), dist_mat=dist_mat)

    def load_arrays(self, arrays: dict, max_vehicle_number: int, check: bool = True,
                    dist_mat: Optional[torch.Tensor] = None) -> None:
        """Same as load_data for instances that are already stacked: fp32 tensors ``coords (I,N,2)``,
        ``demands (I,N)``, ``tw (I,N,2)``, ``service_time (I,)``, ``org_service_horizon (I,)`` on the host
        (pinned memory is copied asynchronously) or on the device.  The tensors are copied, never aliased:
        the window fix-up (env.py:395-409) modifies ``tw`` in place.  ``check=False`` skips the one
        device->host read that turns a collapsed window into the reference's AssertionError."""
        dev = self.device

        def up(name, t):
            # persistent device buffers: the same pointers across loads of equally shaped batches (lets callers
            # replay a captured CUDA graph of the rollout); always a copy, never an alias of the caller's tensor
            cur = getattr(self, name)
            if cur is None or tuple(cur.shape) != tuple(t.shape):
                cur = torch.empty(tuple(t.shape), dtype=torch.float32, device=dev)
                setattr(self, name, cur)
            cur.copy_(t.to(dtype=torch.float32) if t.device != dev else t.to(torch.float32), non_blocking=True)

        up("coords", arrays["coords"])
        up("demands", arrays["demands"])
        up("tw", arrays["tw"])
        up("service_time", arrays["service_time"])
        up("org_service_horizon", arrays["org_service_horizon"])
        I, gs = int(self.coords.size(0)), int(self.coords.size(1))
        kv = int(max_vehicle_number)
        self.n_inst = I
        self.graph_size = gs
        self.max_vehicle_number = int(kv + np.floor(np.log(kv)))     # env.py:362
        if self.vehicle_capacity is None:
            self.vehicle_capacity, self.service_horizon = 1.0, 1.0
        self.k_nbh_size = math.ceil(self.k_nbh_frac * gs)     # graph_utils.py:99
        K = self.max_concurrent_vehicles
        if self.max_vehicle_number > 32767 or gs > 32767:
            raise ValueError("graph too large for the int16 index tables")
        N, k = gs, self.k_nbh_size
        shape_key = (I, N, k)
        if getattr(self, "_table_key", None) != shape_key:
            ta = self._table_alloc = _GuardedAlloc()
            self._dist_mat = ta.zeros((I, N, N), torch.float32, dev, "dist")
            self.time_to_depot = ta.zeros((I, N), torch.float32, dev, "ttd")
            self._knn = ta.zeros((I, N, k), torch.int16, dev, "knn")
            self._knn_w = ta.zeros((I, N, k), torch.float32, dev, "knn_w")
            self.ordered_idx = ta.zeros((I, N), torch.int16, dev, "order")
            self._inst_status = ta.zeros((I,), torch.int32, dev, "inst_status")
            self._static_emb = ta.zeros((I, N, 256), torch.float32, dev, "static_emb")
            self._h0 = ta.zeros((I, N, 128), torch.float32, dev, "h0")
            # column-chunk-major copies read by the single-tile tensor-core step (include/jampr_b200.h)
            self._h0_t = ta.zeros((I, 32, N, 4) if N <= 128 else (I, (N + 127) // 128, 32, 128, 4), torch.float32, dev, "h0_t")
            self._static_emb_t = ta.zeros((I, 64, N, 4), torch.float32, dev, "static_emb_t") if N <= 128 else None
            # packed edge tables of the tensor-core policy step (filled by jampr_node_encoder)
            self._tc_tab = ta.zeros((I, N, (k + 3) // 4 * 4), torch.int32, dev, "tc_tab")
            self._tc_dtab = ta.zeros((I, (N + 7) // 8 * 8), torch.int32, dev, "tc_dtab")
            # scratch of the tensor-core node encoder (csrc/gnn_large.cu): 16-bit ping-pong embedding + fp32 skip
            npad = (N + 127) // 128 * 128
            self._lg_h16 = ta.zeros((2, I, npad, 128), torch.int16, dev, "lg_h16")
            self._lg_skip = ta.zeros((I, npad, 128), torch.float32, dev, "lg_skip")
            self._table_key = shape_key
        self.bs = I * self._num_samples
        self._dims = _cabi.Dims(I, self._num_samples, N, K, k, self.max_vehicle_number, min(64, N),
                                self.tour_graph_update_step)
        p = _cabi.Instances()
        for name, t in (("coords", self.coords), ("demand", self.demands), ("tw", self.tw),
                        ("service", self.service_time), ("horizon", self.org_service_horizon),
                        ("dist", self._dist_mat), ("ttd", self.time_to_depot), ("knn", self._knn),
                        ("knn_w", self._knn_w), ("order", self.ordered_idx), ("inst_status", self._inst_status),
                        ("static_emb", self._static_emb), ("h0", self._h0), ("tc_tab", self._tc_tab),
                        ("tc_dtab", self._tc_dtab), ("lg_h16", self._lg_h16), ("lg_skip", self._lg_skip),
                        ("h0_t", self._h0_t), ("static_emb_t", self._static_emb_t)):
            setattr(p, name, _cabi.ptr(t))
        given = None
        if dist_mat is not None:
            given = dist_mat.to(device=dev, dtype=torch.float32).contiguous()
            assert tuple(given.shape) == (I, N, N)
        p.dist_given = _cabi.ptr(given)
        self._inst = p
        tw_in = None if self.inference else self.tw.clone()
        _cabi.check(self._lib.jampr_setup_instances(self._dims, self._inst, self._stream()), "jampr_setup_instances")
        p.dist_given = None           # consumed by the kernel enqueued above (same stream keeps `given` alive)
        self._dist_given_keepalive = given
        if tw_in is not None:
            self.tw.copy_(tw_in)      # the window fix-up is an inference-only step in the reference (env.py:395)
        if self.inference and check:
            # reference asserts on the shrunk windows (env.py:408)
            if bool((self._inst_status & _cabi.ST_TW_COLLAPSE).any().item()):
                raise AssertionError("time window fix for the return to the depot collapses a window (env.py:408)")
        self._static_version += 1
        self._has_instance = True
        self._is_reset = False

    # ------------------------------------------------------------------ reset (env.py:146-201, 1221-1297)
    def _pomo_start_nodes(self) -> torch.Tensor:
        """Start nodes of the POMO pseudo-steps (env.py:1221-1270) on the device (jampr_pomo_start_nodes): the depot
        neighbourhood without the depot, reduced to the earliest window starts, K distinct uniform picks per rollout
        — or the deterministic single start node per sample.  Every reset() since the last seed() uses a fresh stream."""
        P, K, k = self._num_samples, self.max_concurrent_vehicles, self.k_nbh_size
        if self.pomo_single_start_node:
            if P > k - 1:
                raise RuntimeError(f"sample set for POMO has size {k - 1} which is too small for {P} samples!")
            kk, n_start = k - 1, 1
        else:
            kk, n_start = math.floor((k - 1) * self.pomo_nbh_frac), K
            if kk < K:
                raise RuntimeError(f"sample set for POMO has size {kk} which is too small for {P} samples! "
                                   f"Try increasing 'k_nbh_frac' of env.")
        if k - 1 > 256:
            raise NotImplementedError("POMO start selection supports neighbourhoods of up to 257 nodes")
        out = torch.empty(self.bs, n_start, dtype=torch.int32, device=self.device)
        seed = (int(self._seed_value) * 0x9E3779B1 + self._n_resets * 0x85EBCA77 + 0x5bd1e995) & 0xFFFFFFFFFFFFFFFF
        self._n_resets += 1
        _cabi.check(self._lib.jampr_pomo_start_nodes(self._dims, self._inst, kk, n_start,
                                                     int(self.pomo_single_start_node), seed, _cabi.ptr(out),
                                                     self._stream()), "jampr_pomo_start_nodes")
        return out

    def reset(self, start_nodes: Optional[torch.Tensor] = None) -> RPObs:
        """Reset the simulator and return the initial observation.  ``start_nodes`` (bs, n_start) overrides
        the random POMO draw (used to replay a recorded episode)."""
        assert self._has_instance, "need to load instance first."
        d = self._dims
        if (self.state is None or self.state.bs != self.bs or self.state.N != d.n_nodes or self.state.k != d.k_nbh
                or self.state.K_max != d.k_max):
            self.state = DeviceState(self.bs, d.n_nodes, d.n_slots, d.k_nbh, d.k_max, d.plan_len, self.device)
        self._step = 0
        n_start = 0
        sn = None
        if self.pomo:
            sn = self._pomo_start_nodes() if start_nodes is None else \
                start_nodes.to(self.device, torch.int32).reshape(self.bs, -1).contiguous()
            n_start = sn.size(1)
        _cabi.check(self._lib.jampr_env_reset(d, self._inst, self.state.struct(), _cabi.ptr(sn), n_start,
                                              self._stream()), "jampr_env_reset")
        self._start_nodes = sn
        self._step += n_start
        self._graph_fresh = True
        self._is_reset = True
        self._done = False
        return self._get_observation()

    # ------------------------------------------------------------------ step (env.py:203-311)
    def step(self, action: torch.Tensor):
        """One environment step.  Returns (obs | None, cost (bs,), done, info) like the reference."""
        assert self._is_reset
        assert action.size(0) == self.bs and action.size(1) == 2
        st = self.state
        act = action.to(device=self.device, dtype=torch.int32).contiguous()
        _cabi.check(self._lib.jampr_env_step(self._dims, self._inst, st.struct(), _cabi.ptr(act), self._step,
                                             self._stream()), "jampr_env_step")
        # the reference returns `done` as a Python bool: one device->host read per step (env.py:281);
        # the fused rollout (jampr_rollout) has no such read
        done = int(st["ctl"][_cabi.CTL_DONE_STEP].item()) >= 0
        self._raise_status()
        self._graph_fresh = (self._step <= self.max_concurrent_vehicles + 1) or \
                            (self._step % self.tour_graph_update_step == 0)      # env.py:801
        cost = st["cost"].clone()
        info = {
            "current_total_cost": st["total"].cpu().numpy(),
            "k_used": self.k_used.cpu().numpy() if done else [-1],
            "max_tour_len": int((st["plan"] == 0).to(torch.uint8).argmax(dim=-1).max().item()),
        }
        self._step += 1
        if done:
            self._has_instance = False
            self._is_reset = False
            self._done = True
        return (self._get_observation() if not done else None), cost, done, info

    def _raise_status(self, allow_warn: bool = True):
        """Turn the per-rollout status words into the reference's exceptions / warnings."""
        status = self.state["status"]
        if int(status.max().item()) == 0:
            return
        has = lambda bit: bool(((status & bit) != 0).any().item())
        if has(_cabi.ST_TOUR_OVERFLOW):
            n_inf = int(((status & _cabi.ST_TOUR_OVERFLOW) != 0).sum().item())
            raise RuntimeError(f"Current rollout could not solve at least {n_inf} instances.")
        if has(_cabi.ST_NO_FEASIBLE):
            raise RuntimeError("no feasible nodes left: %s" %
                               ((status & _cabi.ST_NO_FEASIBLE) != 0).nonzero().view(-1).tolist()[:8])
        if has(_cabi.ST_NAN_LOGITS):
            raise AssertionError("Logits contain NANs!")
        if allow_warn and has(_cabi.ST_FLEET_EXHAUSTED):
            warnings.warn("used all available vehicles!", RuntimeWarning)

    # ------------------------------------------------------------------ observation (env.py:1042-1073)
    def _get_observation(self) -> RPObs:
        st = self.state
        full = self.debug_lvl > 0
        node_features = tour_plan = tour_features = None
        if full:
            node_features = self.node_features()
            tour_plan = st["plan"].gather(1, st["row"].long()[:, :, None].expand(self.bs, st.K, st.L)).long()
            tour_features = self.tour_features()
        return RPObs(
            batch_size=self.bs,
            node_features=node_features,
            node_nbh_edges=None,
            node_nbh_weights=None,
            tour_plan=tour_plan,
            tour_features=tour_features,
            tour_edges=None,
            tour_weights=None,
            nbh=st["nbh"],
            nbh_mask=st["mask"].view(torch.bool) if st["mask"].dtype == torch.uint8 else st["mask"],
            state=self,
        )

    def node_features(self) -> torch.Tensor:
        """(n_inst, N, 7): coords, demand, tw, service time, time to depot (env.py:1047-1053)."""
        I, N = self.n_inst, self.graph_size
        return torch.cat((self.coords, self.demands[:, :, None], self.tw,
                          self.service_time[:, None, None].expand(I, N, 1), self.time_to_depot[:, :, None]), -1)

    def tour_features(self) -> torch.Tensor:
        """(bs, K, 5) as env.py:1061-1068."""
        st = self.state
        num = (self.max_vehicle_number - st["row"]).float()
        # tensor / tensor: ATen turns division by a host scalar into a multiplication by the reciprocal on CUDA
        return torch.stack((st["cur"].float(), st["cap"], st["time"], st["cttd"],
                            num / torch.full_like(num, float(self.max_vehicle_number))), -1)

    # ------------------------------------------------------------------ properties (env.py:734-767)
    @property
    def num_samples(self):
        return self._num_samples

    @property
    def depot_node(self) -> torch.Tensor:
        return torch.zeros(self.bs, dtype=torch.long, device=self.device)

    @property
    def visited(self) -> torch.Tensor:
        return self.state["visited"][:, 1:].bool()

    @property
    def _visited(self) -> torch.Tensor:
        return self.state["visited"].bool()

    @property
    def tour_plan(self) -> torch.Tensor:
        return self.state["plan"]

    @property
    def _finished(self) -> torch.Tensor:
        return (self.state["vflags"] & 2) != 0

    @property
    def active_vehicles(self) -> torch.Tensor:
        return (self.state["vflags"] & 1) != 0

    @property
    def active_to_plan_idx(self) -> torch.Tensor:
        return self.state["row"].long()

    @property
    def cur_node(self) -> torch.Tensor:
        return self.state["cur"].long()

    @property
    def k_used(self) -> torch.Tensor:
        """Number of vehicles used per rollout (env.py:754-759)."""
        st = self.state
        used = self._finished.clone()
        en_route = st["cur"] != 0
        rows = st["row"].long()
        used.scatter_(1, rows, used.gather(1, rows) | en_route)
        return used.sum(-1)

    @property
    def total_cost(self) -> torch.Tensor:
        return self.state["total"].clone()

    def true_cost(self) -> torch.Tensor:
        """Batch-independent cost recomputed from the tour buffers."""
        out = torch.empty(self.bs, dtype=torch.float32, device=self.device)
        _cabi.check(self._lib.jampr_true_cost(self._dims, self._inst, self.state.struct(), _cabi.ptr(out),
                                              self._stream()), "jampr_true_cost")
        return out

    def gather_best(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Best-of-samples per instance (training.py:80-94): (cost (I,), sample idx (I,), tours (I,K_max,L))."""
        d = self._dims
        cost = torch.empty(d.n_inst, dtype=torch.float32, device=self.device)
        idx = torch.empty(d.n_inst, dtype=torch.int32, device=self.device)
        plan = torch.empty(d.n_inst, d.k_max, d.plan_len, dtype=torch.int16, device=self.device)
        _cabi.check(self._lib.jampr_gather_best(d, self.state.struct(), _cabi.ptr(cost), _cabi.ptr(idx),
                                                _cabi.ptr(plan), self._stream()), "jampr_gather_best")
        return cost, idx, plan

    # ------------------------------------------------------------------ solution I/O (env.py:465-718)
    def export_sol(self, num_best: int = 3, num_other: int = 3, mode: str = "random"):
        """Export tour plans as nested lists (env.py:465-573); ``mode`` = "random" or "similarity" (host-side
        difflib matching like the reference, jampr_plus_b200/selection.py)."""
        P = self._num_samples
        st = self.state
        plan = st["plan"].view(-1, P, st.K_max, st.L)
        total = st["total"].view(-1, P)
        if P > 1:
            n_smp = num_best + num_other
            assert P > n_smp, (f"specified selection of total of {n_smp} samples but env was configured "
                               f"with num_samples = {P} < {n_smp}!")
            cost_idx = total.sort(dim=-1, stable=True).indices
            idx = cost_idx[:, :num_best]
            if mode == "similarity":
                # host-side difflib matching like the reference (env.py:498-545, seq_match.py)
                from ..selection import select_similarity
                all_tours = self._sol_to_list(plan)
                all_costs = total.cpu().tolist()
                tours, costs = [], []
                for bst, smps, cst in zip(idx.cpu().tolist(), all_tours, all_costs):
                    t, c = select_similarity(smps, cst, bst, num_other)
                    tours.append(t)
                    costs.append(c)
                return tours, costs
            if mode != "random":
                raise ValueError(f"unknown selection mode: {mode}.")
            if num_other > 0:
                w = torch.ones(plan.size(0), P - num_best, device=self.device)
                rnd = torch.multinomial(w, num_other, replacement=False, generator=self._gen)
                idx = torch.cat((idx, cost_idx[:, num_best:].gather(-1, rnd)), -1)
            tp = plan.gather(1, idx[:, :, None, None].expand(-1, -1, st.K_max, st.L))
            return self._sol_to_list(tp), total.gather(-1, idx).cpu().tolist()
        return self._sol_to_list(plan), st["total"].cpu().tolist()

    def _sol_to_list(self, tp: torch.Tensor):
        tp = tp.cpu().numpy()
        full = set(range(self.graph_size))
        tours = []
        for plans in tp:
            tour_set = []
            for plan in plans:
                # unvisited customers become singleton tours in the iteration order of the reference's own expression
                # (a Python set difference, env.py:563-570): irrelevant for import_sol, kept for list-level equality
                unq = np.unique(plan)
                singles = [[e] for e in full - set(unq.tolist())] if len(unq) != self.graph_size else []
                rows = plan[plan.sum(-1) > 0]
                tour_set.append([r[r > 0].tolist() for r in rows] + singles)
            tours.append(tour_set)
        return tours

    # ------------------------------------------------------------------ dense solution format
    def _lists_to_dense(self, sol: List[List]) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """Nested lists (instances x samples x tours, the reference's format, env.py:578-581) -> dense arrays
        ``tours (n_sol, T, L) int16`` (customers, zero padded) and ``kind (n_sol, T) uint8`` (1 complete = list starts
        and ends at the depot, 2 partial, 0 none); singleton tours are skipped like the reference skips them."""
        L = self.state.L
        flat = [smp for inst_sol in sol for smp in inst_sol]
        T = max(1, max(sum(1 for t in smp if len(t) > 1) for smp in flat))
        tours = np.zeros((len(flat), T, L), dtype=np.int16)
        kind = np.zeros((len(flat), T), dtype=np.uint8)
        for j, smp in enumerate(flat):
            t_idx = 0
            for tour in smp:
                l = len(tour)
                if l <= 1:
                    continue
                if tour[0] == tour[-1] == 0:
                    body, kd = tour[1:-1], 1
                else:
                    body, kd = tour, 2
                    if 0 in body:
                        raise ValueError("a partial tour must not contain the depot")
                if len(body) > L:
                    raise RuntimeError(f"tour with {len(body)} customers exceeds the tour buffer ({L})")
                tours[j, t_idx, :len(body)] = body
                kind[j, t_idx] = kd
                t_idx += 1
        return torch.from_numpy(tours), torch.from_numpy(kind), T

    def _import_dense(self, tours: torch.Tensor, kind: torch.Tensor, src_index: Optional[torch.Tensor],
                      total_in: Optional[torch.Tensor]) -> RPObs:
        """import_sol on the device (jampr_env_import_tours + jampr_env_import_finish)."""
        st = self.state
        dev = self.device
        tours = tours.to(dev, torch.int16).contiguous()
        kind = kind.to(dev, torch.uint8).contiguous()
        n_sol, T, Lt = tours.shape
        if src_index is not None:
            src_index = src_index.to(dev, torch.int32).contiguous()
            assert src_index.numel() == self.bs
        else:
            assert n_sol == self.bs
        if total_in is not None:
            total_in = total_in.to(dev, torch.float32).contiguous()
            assert total_in.numel() == self.bs
        if self._max_step_buf is None:
            self._max_step_buf = torch.zeros(1, dtype=torch.int32, device=dev)
        _cabi.check(self._lib.jampr_env_import_tours(self._dims, self._inst, st.struct(), _cabi.ptr(tours), _cabi.ptr(kind),
                                                     T, Lt, _cabi.ptr(src_index), _cabi.ptr(total_in), self._stream()),
                    "jampr_env_import_tours")
        _cabi.check(self._lib.jampr_env_import_finish(self._dims, self._inst, st.struct(), _cabi.ptr(self._max_step_buf),
                                                      self._stream()), "jampr_env_import_finish")
        # the reference rebuilds the tour graph with the PREVIOUS episode's step counter (env.py:700 before
        # :704): same rule, same stale value
        prev_step = self._step if self._step is not None else 0
        self._graph_fresh = (prev_step <= st.K + 1) or (prev_step % self.tour_graph_update_step == 0)
        self._has_instance = True
        self._is_reset = True
        self._done = False
        self._step = int(self._max_step_buf.item())
        self._raise_status()
        return self._get_observation()

    def import_sol(self, sol: List[List], cost: Optional[List] = None) -> RPObs:
        """Import partial solutions (env.py:575-706).  The nested lists are flattened to dense tour arrays on the host
        (no per-node work), everything else — tour buffers, vehicle slots, capacity / clock of the partial tours,
        recomputed totals (env.py:708-718), visited flags, tour links, first observation — runs on the device."""
        assert self.inference, "Can only import solutions in inference mode."
        assert self.state is not None, "import_sol needs a previous reset() (reference reuses its buffers)"
        P = self._num_samples
        I = self.n_inst
        assert len(sol) == I
        n_samples = np.array([len(s) for s in sol])
        assert np.all(n_samples == n_samples[0])
        n_samples = int(n_samples[0])
        if self.pomo:
            assert P == n_samples, (f"imported solutions for POMO must include num_samples samples, "
                                    f"but got {n_samples} != {P}")
        else:
            assert P % n_samples == 0
        fact = P // n_samples
        tours, kind, _ = self._lists_to_dense(sol)
        if int(kind.ne(0).sum(-1).max()) > self.state.K_max:
            raise RuntimeError("Number of tours of provided solution is larger than max_num_vehicles!")
        src = None
        if fact > 1:                    # _expand_sample_dimension (env.py:1204-1219)
            src = torch.arange(I * n_samples, dtype=torch.int32).repeat_interleave(fact)
        total_in = None
        if cost is not None:
            c = torch.as_tensor(np.asarray(cost, dtype=np.float32).reshape(-1))
            total_in = c if c.numel() == self.bs else c.repeat_interleave(fact)[: self.bs]
        return self._import_dense(tours, kind, src, total_in)

    def _import_sol_host(self, sol: List[List], cost: Optional[List] = None) -> RPObs:
        """Round-1 implementation of import_sol, kept as the CHECKER of the device path (tests compare the state arrays
        of both bit for bit): the reference's triple loop (env.py:612-690) in numpy on the host, state arrays uploaded
        once, then jampr_env_import_finish."""
        assert self.inference, "Can only import solutions in inference mode."
        assert self.state is not None, "import_sol needs a previous reset() (reference reuses its buffers)"
        P = self._num_samples
        I = self.n_inst
        assert len(sol) == I
        n_samples = np.array([len(s) for s in sol])
        assert np.all(n_samples == n_samples[0])
        n_samples = int(n_samples[0])
        if self.pomo:
            assert P == n_samples, (f"imported solutions for POMO must include num_samples samples, "
                                    f"but got {n_samples} != {P}")
        else:
            assert P % n_samples == 0
        fact = P // n_samples
        st = self.state
        bs, K, K_max, L = self.bs, st.K, st.K_max, st.L
        plan = np.zeros((bs, K_max, L), dtype=np.int16)
        vfl = np.zeros((bs, K_max), dtype=np.uint8)
        row = np.zeros((bs, K), dtype=np.int32)
        nxt = np.zeros((bs, K), dtype=np.int32)
        cur = np.zeros((bs, K), dtype=np.int32)
        cap = np.ones((bs, K), dtype=np.float32)
        tim = np.zeros((bs, K), dtype=np.float32)
        cttd = np.zeros((bs, K), dtype=np.float32)
        total = np.zeros(bs, dtype=np.float32)
        dist = self._dist_mat.cpu().numpy()
        tw0 = self.tw[:, :, 0].cpu().numpy()
        dem = self.demands.cpu().numpy()
        ttd = self.time_to_depot.cpu().numpy()
        srv = self.service_time.cpu().numpy()

        def recompute(tour, i):      # env.py:708-718, fp32 accumulation like the reference tensors
            tm = np.float32(0.0)
            prev = 0
            for n in tour:
                tm = np.float32(tm + dist[i, prev, n])
                tm = np.float32(tm + np.float32(np.maximum(np.float32(tw0[i, n] - tm), np.float32(0)) + srv[i]))
                prev = n
            return float(tm)

        b = 0
        for i, inst_sol in enumerate(sol):
            for smp in inst_sol:
                num_partial = 0
                t_idx = 0
                costs = []
                for tour in smp:
                    l = len(tour)
                    if l <= 1:
                        continue
                    if t_idx >= K_max:
                        raise RuntimeError("Number of tours of provided solution is larger than max_num_vehicles!")
                    if tour[0] == tour[-1] == 0:
                        plan[b, t_idx, :l - 1] = np.asarray(tour[1:], dtype=np.int16)
                        vfl[b, t_idx] |= 2
                        costs.append(recompute(tour, i))
                    else:
                        assert num_partial < K
                        t = np.asarray(tour, dtype=np.int64)
                        plan[b, t_idx, :l] = t.astype(np.int16)
                        cur[b, num_partial] = t[-1]
                        cap[b, num_partial] = np.float32(1.0) - dem[i, t].sum(dtype=np.float32)
                        cttd[b, num_partial] = ttd[i, t[-1]]
                        tm = recompute(tour, i)
                        costs.append(tm)
                        tim[b, num_partial] = tm
                        nxt[b, num_partial] = l
                        vfl[b, t_idx] |= 1
                        row[b, num_partial] = t_idx
                        num_partial += 1
                    t_idx += 1
                assert num_partial <= K
                for s in range(num_partial, K):
                    free = np.flatnonzero(vfl[b] == 0)
                    nf = int(free[0]) if free.size else 0
                    vfl[b, nf] |= 1
                    row[b, s] = nf
                # rows were handed out in ascending order, which is what the reference's
                # active_vehicles.nonzero() re-derivation yields (env.py:697)
                total[b] = np.float32(sum(costs)) if cost is None else 0.0
                for r in range(1, fact):      # _expand_sample_dimension (env.py:1204-1219)
                    for arr in (plan, vfl, row, nxt, cur, cap, tim, cttd, total):
                        arr[b + r] = arr[b]
                b += fact
        if cost is not None:
            total = np.repeat(np.asarray(cost, dtype=np.float32).reshape(-1), fact)[:bs] \
                if np.asarray(cost).size != bs else np.asarray(cost, dtype=np.float32).reshape(-1)
        for name, arr in (("plan", plan), ("vflags", vfl), ("row", row), ("nxt", nxt), ("cur", cur), ("cap", cap),
                          ("time", tim), ("cttd", cttd), ("total", total)):
            st[name].copy_(torch.from_numpy(np.ascontiguousarray(arr)).to(self.device))
        st["row_last"].zero_()
        st["cost"].zero_()
        st["status"].zero_()
        for name in ("visited", "pred", "succ"):
            st[name].zero_()
        max_step = torch.zeros(1, dtype=torch.int32, device=self.device)
        _cabi.check(self._lib.jampr_env_import_finish(self._dims, self._inst, st.struct(), _cabi.ptr(max_step),
                                                      self._stream()), "jampr_env_import_finish")
        # the reference rebuilds the tour graph with the PREVIOUS episode's step counter (env.py:700 before
        # :704): same rule, same stale value
        prev_step = self._step if self._step is not None else 0
        self._graph_fresh = (prev_step <= K + 1) or (prev_step % self.tour_graph_update_step == 0)
        self._has_instance = True
        self._is_reset = True
        self._done = False
        self._step = int(max_step.item())
        self._raise_status()
        return self._get_observation()

    def local_search(self, max_iters: int = 200, update_totals: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """Improve the finished solutions held by this env (tour buffers of all rollouts) by the device local search
        (relocate / 2-opt*, csrc/local_search.cu) — the stand-in for the reference's OR-tools step (lns.py `_local_search`,
        gort.py:129-290).  With ``update_totals`` the reported total of every rollout becomes the batch-independent cost
        of its solution (served tours + 2 * time_to_depot per unserved customer), so that improved and untouched
        solutions are compared on one scale.  Returns (cost of the served tours (bs,), moves applied (bs,))."""
        assert self.state is not None
        st = self.state
        cost = torch.empty(self.bs, dtype=torch.float32, device=self.device)
        moves = torch.empty(self.bs, dtype=torch.int32, device=self.device)
        _cabi.check(self._lib.jampr_local_search(self._dims, self._inst, _cabi.ptr(st["plan"]), self.bs, None,
                                                 int(max_iters), _cabi.ptr(cost), _cabi.ptr(moves), self._stream()),
                    "jampr_local_search")
        if update_totals:
            st["total"].copy_(self.true_cost())
        return cost, moves

    def destruct(self, num_best: int = 2, num_other: int = 3, num_partial: Optional[int] = None,
                 rm_partial_mode: str = "random", rm_complete_mode: str = "random",
                 frac_rm_nodes_from_partial: float = 0.5, frac_rm_complete_routes: float = 0.15,
                 seed: int = 1, iteration: int = 0, incumbent_plan: Optional[torch.Tensor] = None,
                 incumbent_cost: Optional[torch.Tensor] = None, sources: Optional[torch.Tensor] = None,
                 **kwargs) -> RPObs:
        """Tensor-native select -> bootstrap -> destroy -> import on the finished episode held by this env: the
        operator the reference left as a stub (env.py:720-723), for its ``random_from_route`` destruction
        (destructor.py:86-205) and the selection / bootstrap of lns.py:296-313 with ``mode="random"``.

        Per instance the ``num_best`` cheapest samples and ``num_other`` random others are selected (export_sol,
        env.py:484-497, 546-552); if ``incumbent_plan (I, K_max, L)`` / ``incumbent_cost (I,)`` are given and the
        incumbent is not among them it replaces the most expensive one (lns.py:209-222); the selection is repeated
        to ``num_samples`` rollouts, each is mutilated with its own random stream and imported.  ``sources``
        ``(I, n_sel, K_max, L)`` replaces the selection (sharded LNS: the globally selected tour buffers).
        Nothing leaves the device except the new step counter.  Returns the first observation of the repair episode."""
        assert self.inference and self.state is not None
        st = self.state
        I, P, R, L = self.n_inst, self._num_samples, st.K_max, st.L
        n_sel = num_best + num_other
        if sources is None:
            assert P > n_sel and P % n_sel == 0, "num_samples must be a multiple of num_best + num_other (lns.py:307)"
            plan = st["plan"].view(I, P, R, L)
            total = st["total"].view(I, P)
            order = total.sort(dim=-1, stable=True).indices
            idx = order[:, :num_best]
            if num_other > 0:
                w = torch.ones(I, P - num_best, device=self.device)
                rnd = torch.multinomial(w, num_other, replacement=False, generator=self._gen)
                idx = torch.cat((idx, order[:, num_best:].gather(-1, rnd)), -1)
            sources = plan.gather(1, idx[:, :, None, None].expand(-1, -1, R, L)).contiguous()
            sel_cost = total.gather(-1, idx)
            if incumbent_plan is not None:
                inc = incumbent_cost.to(self.device, torch.float32).view(I, 1)
                worst = sel_cost.argmax(-1)
                absent = ~(sel_cost == inc).any(-1)
                rows = torch.nonzero(absent).view(-1)
                sources[rows, worst[rows]] = incumbent_plan.to(self.device, torch.int16)[rows]
        else:
            sources = sources.to(self.device, torch.int16).contiguous()
            n_sel = sources.size(1)
            assert P % n_sel == 0
        # bootstrap (lns.py:304-310): the list of selected solutions repeated P / n_sel times
        src_index = (torch.arange(P, device=self.device, dtype=torch.int32) % n_sel)[None, :] \
            + (torch.arange(I, device=self.device, dtype=torch.int32) * n_sel)[:, None]
        src_index = src_index.reshape(-1).contiguous()
        prm = _cabi.Destroy()
        prm.num_partial = int(self.max_concurrent_vehicles if num_partial is None else num_partial)
        prm.rm_partial_mode = _cabi.RM_PARTIAL_MODES[rm_partial_mode]
        prm.rm_complete_mode = _cabi.RM_COMPLETE_MODES[rm_complete_mode]
        prm.frac_partial = float(frac_rm_nodes_from_partial)
        prm.frac_complete = float(frac_rm_complete_routes)
        prm.iteration = int(iteration)
        prm.seed = int(seed)
        if self._destroy_buf is None or self._destroy_buf[0].size(0) != self.bs:
            self._destroy_buf = (torch.zeros(self.bs, R, L, dtype=torch.int16, device=self.device),
                                 torch.zeros(self.bs, R, dtype=torch.uint8, device=self.device))
        tours, kind = self._destroy_buf
        import ctypes as _C
        _cabi.check(self._lib.jampr_lns_destroy(self._dims, _cabi.ptr(sources.view(-1, R, L)), _cabi.ptr(src_index),
                                                _C.byref(prm), _cabi.ptr(tours), _cabi.ptr(kind), self._stream()),
                    "jampr_lns_destroy")
        self._last_sources = sources
        return self._import_dense(tours, kind, None, None)
