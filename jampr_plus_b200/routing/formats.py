"""Data contract between environment and policy.

Field names/order mirror the reference tuples so that reference callers can switch
without edits: ``RPInstance`` = reference ``lib/routing/formats.py:26-60``,
``RPObs`` = ``lib/routing/formats.py:63-74``.
"""
from typing import Any, List, NamedTuple, Union

import numpy as np
import torch

__all__ = ["RPInstance", "RPObs"]

_Array = Union[np.ndarray, torch.Tensor]
_Scalar = Union[float, int]


class RPInstance(NamedTuple):
    coords: _Array                           # (N, 2) in [0, 1]^2, row 0 = depot
    demands: _Array                          # (N,) normalised by vehicle capacity
    tw: _Array                               # (N, 2) normalised by service horizon
    service_time: Union[_Array, float]       # one value per instance (all customers share it)
    graph_size: int                          # N, depot included
    org_service_horizon: _Scalar             # depot due date in original units (cost un-normalisation)
    max_vehicle_number: int                  # fleet size kv; the env uses K_max = kv + floor(ln kv)
    vehicle_capacity: float = 1.0            # always 1 after normalisation
    service_horizon: float = 1.0             # always 1 after normalisation
    depot_idx: List = [0]                    # the depot is node 0
    type: Union[int, str] = ""               # Solomon type "1" / "2" (selects the checkpoint family)
    tw_frac: Union[float, str] = ""          # fraction of customers with a real time window

    def __getitem__(self, key):
        # reference instances are indexed by field name (env.py:1031-1033)
        if isinstance(key, str):
            return getattr(self, key)
        return tuple.__getitem__(self, key)

    def get(self, key, default_val: Any = None):
        try:
            return self[key]
        except (AttributeError, IndexError):
            return default_val


class RPObs(NamedTuple):
    batch_size: int                   # number of rollouts (instances x samples)
    node_features: torch.Tensor       # (bs, N, 7) f32
    node_nbh_edges: torch.Tensor      # (2, bs*(N + (N-1)k)) i64 running index
    node_nbh_weights: torch.Tensor    # (bs*(N + (N-1)k),) f32
    tour_plan: torch.Tensor           # (bs, K, L) i64 - rows of the K active vehicles
    tour_features: torch.Tensor       # (bs, K, 5) f32
    tour_edges: torch.Tensor          # (2, E_t) i64 or empty(0) on skip steps
    tour_weights: torch.Tensor        # (E_t,) f32
    nbh: torch.Tensor                 # (bs, K, k) i64
    nbh_mask: torch.Tensor            # (bs, K, k) bool, True = infeasible
    state: Any = None                 # extension: the env whose device state the policy kernels read
