"""The LNS configuration loader (jampr_plus_b200/config): defaults equal the reference's composed config where the
mount is available, overrides in hydra's key=value syntax, unknown keys fail loudly."""
import os

import pytest
import yaml

from jampr_plus_b200.config import load_config, select_checkpoint

REF = "/root/reference/config_lns"


def test_overrides_and_errors(tmp_path):
    cfg = load_config(None, ["time_limit=60", "num_samples=40", "ds_cfg.rm_partial_mode=random_cut", "data_path=x.txt"])
    assert cfg.time_limit == 60 and cfg.num_samples == 40 and cfg.ds_cfg.rm_partial_mode == "random_cut"
    assert cfg.data_path == "x.txt" and cfg.env_cfg.tour_graph_update_step == 2
    with pytest.raises(KeyError):
        load_config(None, ["no_such_key=1"])
    with pytest.raises(KeyError):
        load_config(None, ["ds_cfg.nope=1"])
    f = tmp_path / "c.yaml"
    f.write_text("num_best: 4\nds_cfg:\n  frac_rm_complete_routes: 0.3\n")
    cfg = load_config(str(f), ["num_best=5"])
    assert cfg.num_best == 5 and cfg.ds_cfg.frac_rm_complete_routes == 0.3 and cfg.ds_cfg.rm_complete_mode == "random"
    f.write_text("bogus: 1\n")
    with pytest.raises(KeyError):
        load_config(str(f))


def test_checkpoint_selection():
    cfg = load_config()
    assert select_checkpoint(cfg, "1", "0.5") == "checkpoints/pg_pomo_all_1_low_50.ckpt"        # c103 (v1.yaml:4-9)
    assert select_checkpoint(cfg, "1", "1.0") == "checkpoints/pg_pomo_all_1_high_50.ckpt"       # r101
    assert select_checkpoint(cfg, "2", "0.73") == "checkpoints/pg_pomo_all_2_high_50.ckpt"
    cfg = load_config(None, ["checkpoint=my.ckpt"])
    assert select_checkpoint(cfg, "1", "0.5") == "my.ckpt"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference mount not available")
def test_defaults_equal_the_reference_config():
    cfg = load_config()
    search = yaml.safe_load(open(os.path.join(REF, "search", "default.yaml")))
    for k in ("seed", "num_steps", "num_samples", "num_best", "num_other", "selection_mode", "pomo_inference",
              "pomo_single_start_node", "add_best", "time_limit"):
        assert cfg[k] == search[k], k
    assert dict(cfg["ds_cfg"]) == search["ds_cfg"]
    k1 = yaml.safe_load(open(os.path.join(REF, "env_cfg_overrides", "K1.yaml")))["env_cfg"]
    assert cfg.env_cfg.k_nbh_frac == k1["k_nbh_frac"] and cfg.env_cfg.tour_graph_update_step == k1["tour_graph_update_step"]
    v1 = yaml.safe_load(open(os.path.join(REF, "model", "v1.yaml")))["type_twf_map_model"]
    assert {int(t): {float(f): p for f, p in m.items()} for t, m in cfg["type_twf_map_model"].items()} == \
           {int(t): {float(f): p for f, p in m.items()} for t, m in v1.items()}
