"""Minimal configuration loader for the LNS command (hydra / omegaconf are not needed): one YAML file with the
reference's keys (config_lns/search/default.yaml:14-39 and the groups composed by config_lns/config.yaml) plus
``key=value`` / ``group.key=value`` overrides in hydra's command-line syntax."""
import copy
import os
from typing import Any, Dict, Iterable, Optional

import yaml

DEFAULT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lns_default.yaml")


class Cfg(dict):
    """dict with attribute access (what the reference code reads from a DictConfig)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def __setattr__(self, k, v):
        self[k] = v


def _set(cfg: Dict[str, Any], dotted: str, value: Any):
    keys = dotted.split(".")
    cur = cfg
    for k in keys[:-1]:
        if k not in cur or not isinstance(cur[k], dict):
            raise KeyError(f"unknown config group '{k}' in override '{dotted}'")
        cur = cur[k]
    if keys[-1] not in cur:
        raise KeyError(f"unknown config key '{dotted}' (overrides may only set existing keys)")
    cur[keys[-1]] = value


def load_config(path: Optional[str] = None, overrides: Iterable[str] = ()) -> Cfg:
    """Defaults <- optional YAML file (same key structure, partial) <- ``a.b=value`` overrides (YAML-typed values)."""
    cfg = yaml.safe_load(open(DEFAULT))
    if path:
        def merge(dst, src, where=""):
            for k, v in (src or {}).items():
                if k not in dst:
                    raise KeyError(f"unknown config key '{where}{k}' in {path}")
                if isinstance(v, dict) and isinstance(dst[k], dict):
                    merge(dst[k], v, f"{where}{k}.")
                else:
                    dst[k] = copy.deepcopy(v)
        merge(cfg, yaml.safe_load(open(path)))
    for ov in overrides:
        if "=" not in ov:
            raise ValueError(f"override '{ov}' is not of the form key=value")
        k, v = ov.split("=", 1)
        _set(cfg, k.strip(), yaml.safe_load(v))
    return Cfg(cfg)


def select_checkpoint(cfg: Cfg, inst_type: str, tw_frac: str) -> str:
    """Checkpoint for an instance type / time-window fraction (reference lib/lns/lns.py `load_instance`: the fraction is
    rounded to the nearest key of type_twf_map_model[type])."""
    if cfg.get("checkpoint"):
        return cfg["checkpoint"]
    table = cfg["type_twf_map_model"][int(inst_type)]
    keys = sorted(float(k) for k in table)
    f = float(tw_frac)
    best = min(keys, key=lambda k: abs(k - f))
    return table[best] if best in table else table[str(best)]
