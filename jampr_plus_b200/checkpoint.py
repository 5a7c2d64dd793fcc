"""Loading the reference's ``.ckpt`` files (``Runner.state_dict``, runner.py:357-367) without its Python environment.

A reference checkpoint is ``torch.save({"cfg": DictConfig, "policy": OrderedDict, "baseline": ..., "optimizer": ...,
"rng": ..., "epoch": int})``; unpickling it needs ``omegaconf`` (for ``cfg``) and, for the ``pg_rollout_*`` files, the
reference's own ``lib.model`` / ``lib.routing`` packages (the rollout baseline pickles a whole Policy and dataset,
pg_baselines.py:280-285).  None of that is needed to run the policy: only ``ckpt["policy"]`` (plain tensors) and a few
plain values of ``cfg``.  ``load_checkpoint`` unpickles with a permissive class resolver: every class that cannot be
imported becomes an inert stand-in that just keeps its pickled state, tensors and containers load normally.

The shipped checkpoints are absent from the reference mount (``.MISSING_LARGE_BLOBS``), so this is exercised on a
synthetic checkpoint with the same structure (tests/test_checkpoint.py); the key schema the policy accepts is pinned
against the reference module (tests/golden/state_dict_schema.json)."""
import importlib
import pickle
from collections import OrderedDict
from typing import Any, Dict, Tuple

import torch

__all__ = ["load_checkpoint", "plain"]


class _Inert:
    """Stand-in for a class whose module is not installed: keeps whatever pickle hands it."""

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def __setstate__(self, state):
        self._state = state

    def __reduce_ex__(self, protocol):          # never re-pickled, but keep pickle happy
        return (_Inert, ())

    # containers restored through append / extend / __setitem__ (pickle's list / dict protocols)
    def append(self, x):
        self.__dict__.setdefault("_items", []).append(x)

    def extend(self, xs):
        self.__dict__.setdefault("_items", []).extend(xs)

    def __setitem__(self, k, v):
        self.__dict__.setdefault("_map", {})[k] = v


def _inert_class(module: str, name: str):
    return type(name, (_Inert,), {"__module__": module})


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            mod = importlib.import_module(module)
            return getattr(mod, name)
        except Exception:                         # missing package (omegaconf, lib.*, hydra, ...) or attribute
            return _inert_class(module, name)


class _PickleModule:
    """pickle look-alike handed to torch.load (it only needs Unpickler / load / __name__)."""
    __name__ = "pickle"
    Unpickler = _Unpickler

    @staticmethod
    def load(f, **kwargs):
        return _Unpickler(f, **kwargs).load()


def plain(obj: Any) -> Any:
    """Best-effort plain-Python view of an inert object tree (omegaconf's DictConfig keeps its content under
    ``_content`` as nodes with ``_val``): dicts / lists / scalars."""
    if isinstance(obj, _Inert):
        st = getattr(obj, "_state", None)
        if isinstance(st, dict):
            if "_content" in st:
                return plain(st["_content"])
            if "_val" in st:
                return plain(st["_val"])
            return {k: plain(v) for k, v in st.items()}
        if "_map" in obj.__dict__:
            return {k: plain(v) for k, v in obj._map.items()}
        if "_items" in obj.__dict__:
            return [plain(v) for v in obj._items]
        return None
    if isinstance(obj, dict):
        return {k: plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [plain(v) for v in obj]
    return obj


def load_checkpoint(path: str, map_location="cpu") -> Tuple["OrderedDict[str, torch.Tensor]", Dict]:
    """Returns (policy state_dict, meta).  ``meta`` = {"cfg": plain view of the stored config (may be partial),
    "epoch": int or None}.  The state_dict goes straight into ``jampr_plus_b200.Policy.load_state_dict``."""
    ckpt = torch.load(path, map_location=map_location, pickle_module=_PickleModule, weights_only=False)
    if not isinstance(ckpt, dict) or "policy" not in ckpt:
        raise KeyError(f"{path}: not a JAMPR+ checkpoint (no 'policy' entry)")
    sd = ckpt["policy"]
    if not all(isinstance(v, torch.Tensor) for v in sd.values()):
        raise TypeError(f"{path}: 'policy' is not a tensor state_dict")
    meta = {"cfg": plain(ckpt.get("cfg")), "epoch": ckpt.get("epoch") if isinstance(ckpt.get("epoch"), int) else None}
    return OrderedDict(sd), meta
