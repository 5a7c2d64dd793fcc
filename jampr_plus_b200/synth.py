"""Seeded synthetic CVRP-TW instances of the Solomon flavour (numpy only).

The reference draws its benchmark instances from ``lib/routing/generator.py`` driven by
``solomon_stats.pkl`` (fitted KDE/gamma models, `generator.py:286-610`).  Neither travels to the
GPU box, so the benchmark and the large-size tests use this self-contained generator, which keeps
the properties the hot path depends on: coordinates in [0,1]^2 (uniform "r", clustered "c", mixed
"rc"), demands normalised by a capacity of 200, a fraction ``tw_frac`` of customers with a
real time window and the rest open, constant service time, every window closing early enough to
return to the depot, ``max_vehicle_number = 25``.  It is NOT a restatement of the reference generator;
parity fixtures that need reference-generated instances store them in ``tests/golden``.
"""
from typing import List, Sequence

import numpy as np

from .routing.formats import RPInstance

# (service horizon, service time, mean tw width) in original Solomon units, type-1 / type-2
_GROUP = {
    ("r", 1): (230.0, 10.0, 30.0), ("c", 1): (1236.0, 90.0, 200.0), ("rc", 1): (240.0, 10.0, 30.0),
    ("r", 2): (1000.0, 10.0, 120.0), ("c", 2): (3390.0, 90.0, 400.0), ("rc", 2): (960.0, 10.0, 120.0),
}


def _dimacs_np(a, b):
    d = 100.0 * (a - b)
    return np.floor(10.0 * np.sqrt((d * d).sum(-1))) / 10.0


def _coords(rng, group: str, n: int) -> np.ndarray:
    if group == "r":
        depot = rng.uniform(0.35, 0.65, size=(1, 2))
        c = np.concatenate((depot, rng.uniform(size=(n, 2))), 0)
    else:
        n_clu = int(rng.integers(6, 11))
        n_c = n if group == "c" else n // 2
        mu = rng.uniform(0.1, 0.9, size=(n_clu, 2))
        sig = rng.uniform(0.02, 0.05, size=(n_clu, 1))
        which = rng.integers(0, n_clu, size=n_c)
        pts = mu[which] + sig[which] * rng.standard_normal((n_c, 2))
        if group == "rc":
            pts = np.concatenate((pts, rng.uniform(size=(n - n_c, 2))), 0)
            pts = pts[rng.permutation(n)]
        depot = rng.uniform(0.35, 0.65, size=(1, 2))
        c = np.clip(np.concatenate((depot, pts), 0), 0.0, 1.0)
    return c.astype(np.float32)


def sample_instance(rng, graph_size: int = 100, group: str = "r", typ: int = 1,
                    tw_frac: float = 0.5) -> RPInstance:
    n = graph_size
    horizon, service, tw_w = _GROUP[(group, int(typ))]
    cap = 200.0 if int(typ) == 1 else 700.0
    coords = _coords(rng, group, n)
    ttd = _dimacs_np(coords[1:].astype(np.float64), coords[0].astype(np.float64)) / horizon
    st = service / horizon
    eps = 1.0 / horizon
    if group == "c":
        dem = rng.choice([10.0, 20.0, 30.0, 40.0], p=[0.45, 0.3, 0.15, 0.1], size=n)
    else:
        dem = np.clip(np.round(rng.gamma(2.2, 7.0, size=n)), 1.0, 41.0)
    dem = dem / cap
    latest = 1.0 - ttd - st - eps                      # latest admissible window end
    assert np.all(latest > ttd + eps), "instance geometry infeasible"
    tw = np.zeros((n, 2))
    tw[:, 1] = latest                                  # customers without a window
    has_tw = np.zeros(n, dtype=bool)
    has_tw[rng.choice(n, size=int(np.ceil(n * tw_frac)), replace=False)] = True
    width = np.clip(rng.normal(tw_w, 0.3 * tw_w, size=n), 0.25 * tw_w, None) / horizon
    centre = rng.uniform(ttd + eps, np.maximum(latest - 0.5 * width, ttd + 2 * eps))
    start = np.clip(centre - 0.5 * width, 0.0, None)
    end = np.minimum(np.maximum(start + width, ttd + eps), latest)
    start = np.minimum(start, end - 0.01)
    start = np.clip(start, 0.0, None)
    tw[has_tw, 0] = start[has_tw]
    tw[has_tw, 1] = end[has_tw]
    tw = np.concatenate((np.array([[0.0, 1.0]]), tw), 0)
    return RPInstance(
        coords=coords, demands=np.concatenate((np.zeros(1), dem)), tw=tw, service_time=st,
        graph_size=n + 1, org_service_horizon=horizon, max_vehicle_number=25,
        vehicle_capacity=1.0, service_horizon=1.0, depot_idx=[0], type=typ, tw_frac=tw_frac,
    )


def sample_instances(num: int, graph_size: int = 100, seed: int = 1234,
                     groups: Sequence[str] = ("r", "c", "rc"), types: Sequence[int] = (1,),
                     tw_fracs: Sequence[float] = (0.25, 0.5)) -> List[RPInstance]:
    """Round-robin over (group, type, tw_frac) like `generator.py:646-656`, one rng per call."""
    rng = np.random.default_rng(seed)
    combos = [(g, t, f) for g in groups for t in types for f in tw_fracs]
    return [sample_instance(rng, graph_size, *combos[i % len(combos)]) for i in range(num)]


def instances_to_arrays(instances: Sequence[RPInstance]) -> dict:
    """Pack a homogeneous list of instances into plain arrays (fixture / C-ABI staging format)."""
    return {
        "coords": np.stack([np.asarray(x.coords, dtype=np.float32) for x in instances]),
        "demands": np.stack([np.asarray(x.demands, dtype=np.float64) for x in instances]),
        "tw": np.stack([np.asarray(x.tw, dtype=np.float64) for x in instances]),
        "service_time": np.asarray([float(x.service_time) for x in instances], dtype=np.float64),
        "org_service_horizon": np.asarray([float(x.org_service_horizon) for x in instances], dtype=np.float64),
        "max_vehicle_number": np.asarray([int(x.max_vehicle_number) for x in instances], dtype=np.int64),
    }


def arrays_to_instances(arr: dict) -> List[RPInstance]:
    out = []
    for i in range(arr["coords"].shape[0]):
        out.append(RPInstance(
            coords=arr["coords"][i], demands=arr["demands"][i], tw=arr["tw"][i],
            service_time=float(arr["service_time"][i]), graph_size=int(arr["coords"].shape[1]),
            org_service_horizon=float(arr["org_service_horizon"][i]),
            max_vehicle_number=int(arr["max_vehicle_number"][i]),
        ))
    return out
