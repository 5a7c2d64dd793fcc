#!/usr/bin/env python
"""Generate the BASELINE.json benchmark workloads with the UNMODIFIED reference generator.

Run in the build container only (needs /root/reference):

    python tests/golden/make_workload.py

  config 2  RPDataset(cfg all_1_low).seed(1234).sample(4096, graph_size=100)   (lib/routing/generator.py:721-781,
            config/env/all_1_low.yaml: groups r/c/rc, types [1], tw_fracs [0.25, 0.5])
  config 3  RPDataset(cfg all_2_high).seed(1234).sample(512, graph_size=100)   (types [2], tw_fracs [0.75, 1.0])

Output: tests/golden/workload_all_1_low_4096.npz / workload_all_2_high_512.npz with the instance arrays exactly as
the reference's RPInstance fields hold them (coords f32; demands, tw narrowed to f32 = what RPEnv.load_data does with
them, env.py:352-366), plus per-instance scalars.  bench.py and the full-size tests read these files; the generator
itself (and its solomon_stats.pkl) cannot travel to the GPU box.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle", "ref_shims"))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import yaml  # noqa: E402

from lib.routing import RPDataset  # noqa: E402  (reference)


def make(cfg_name: str, n: int, out: str, seed: int = 1234, graph_size: int = 100):
    cfg = yaml.safe_load(open(os.path.join(REF, "config", "env", f"{cfg_name}.yaml")))
    scfg = cfg["_data_cfg"]
    print(cfg_name, scfg)
    ds = RPDataset(cfg=scfg, stats_pth=os.path.join(REF, "solomon_stats.pkl"))
    ds.seed(seed)
    data = ds.sample(sample_size=n, graph_size=graph_size)
    assert len(data) == n
    arr = {
        "coords": np.stack([np.asarray(x.coords) for x in data]).astype(np.float32),
        "demands": np.stack([np.asarray(x.demands) for x in data]).astype(np.float32),
        "tw": np.stack([np.asarray(x.tw) for x in data]).astype(np.float32),
        "service_time": np.array([x.service_time for x in data], dtype=np.float32),
        "org_service_horizon": np.array([x.org_service_horizon for x in data], dtype=np.float32),
        "max_vehicle_number": np.array([x.max_vehicle_number for x in data], dtype=np.int64),
        "type": np.array([str(x.type) for x in data]),
        "tw_frac": np.array([str(x.tw_frac) for x in data]),
    }
    # exactness of the narrowing: RPEnv.load_data converts every field to fp32 (env.py:352-366), so the fp32 copy
    # holds exactly what the reference env computes with
    for k in ("demands", "tw"):
        src = np.stack([np.asarray(getattr(x, k)) for x in data])
        assert np.array_equal(src.astype(np.float32), arr[k])
    np.savez_compressed(out, source=f"RPDataset({cfg_name}).seed({seed}).sample({n}, graph_size={graph_size})", **arr)
    print(out, os.path.getsize(out) / 1e6, "MB", {k: v.shape for k, v in arr.items()})


if __name__ == "__main__":
    make("all_1_low", 4096, os.path.join(HERE, "workload_all_1_low_4096.npz"))
    make("all_2_high", 512, os.path.join(HERE, "workload_all_2_high_512.npz"))
