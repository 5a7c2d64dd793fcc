"""Reference-format checkpoints load without omegaconf / the reference's packages (jampr_plus_b200/checkpoint.py).
The synthetic checkpoint mimics runner.py:357-367: a config object of a class from a module that does not exist at
load time, a tensor state_dict, an object-valued baseline entry, optimizer state, rng, epoch."""
import sys
import types

import torch

from jampr_plus_b200.checkpoint import load_checkpoint
from jampr_plus_b200.packing import validate_state_dict
from jampr_plus_b200.weights import random_state_dict


def test_load_reference_style_checkpoint(tmp_path):
    # fabricate the foreign classes, pickle, then make them unimportable again
    oc = types.ModuleType("omegaconf")
    ocd = types.ModuleType("omegaconf.dictconfig")
    lib = types.ModuleType("lib")
    libm = types.ModuleType("lib.model")
    libp = types.ModuleType("lib.model.policy")

    class DictConfig:
        def __init__(self, content):
            self._content = content

        def __getstate__(self):
            return {"_content": self._content, "_metadata": None}

    class Policy:
        def __init__(self):
            self.blob = torch.zeros(3)

    DictConfig.__module__, DictConfig.__qualname__ = "omegaconf.dictconfig", "DictConfig"
    Policy.__module__, Policy.__qualname__ = "lib.model.policy", "Policy"
    ocd.DictConfig, libp.Policy = DictConfig, Policy
    mods = {"omegaconf": oc, "omegaconf.dictconfig": ocd, "lib": lib, "lib.model": libm, "lib.model.policy": libp}
    sys.modules.update(mods)
    sd = random_state_dict(5)
    sd = {k.replace(".conv.lin_rel.", ".conv.lin_l.").replace(".conv.lin_root.", ".conv.lin_r."): v for k, v in sd.items()}
    path = tmp_path / "pg_pomo_like.ckpt"
    try:
        torch.save({"cfg": DictConfig({"env_cfg": {"max_concurrent_vehicles": 3, "k_nbh_frac": 0.25}, "graph_size": 50}),
                    "policy": sd, "baseline": {"policy": Policy()}, "optimizer": {"state": {}, "param_groups": []},
                    "rng": torch.get_rng_state(), "epoch": 49}, path)
    finally:
        for m in mods:
            sys.modules.pop(m, None)
    got, meta = load_checkpoint(str(path))
    assert meta["epoch"] == 49
    assert meta["cfg"]["env_cfg"]["max_concurrent_vehicles"] == 3 and meta["cfg"]["graph_size"] == 50
    canon = validate_state_dict(got)                      # older PyG aliases accepted, schema exact
    ref = random_state_dict(5)
    assert set(canon) == set(ref)
    for k in ref:
        assert torch.equal(canon[k], ref[k])
