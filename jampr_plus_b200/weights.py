"""Checkpoint schema of the JAMPR policy and seeded stand-in weights.

The key names and shapes below are the ``state_dict`` of the reference ``Policy`` built from
``config/policy/gnn_attn.yaml:5-38`` (reference ``lib/model/policy.py:36-95``,
``encoders/node_encoder.py:79-126``, ``encoders/tour_encoder.py:101-157``,
``decoders/attn_decoder.py:49-54``); ``tests/golden/make_golden.py`` checks them against the
reference module in the build container (``tests/golden/state_dict_schema.json``).  The shipped
checkpoints are absent from the reference mount (``.MISSING_LARGE_BLOBS``), so parity and the
benchmark use :func:`random_state_dict` — numpy-seeded, Linear-style uniform init, with
``decoder.out_proj.weight`` scaled by 0.1 so the pointer logits do not saturate the 10*tanh clip
(SURVEY.md §0.7).
"""
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch

H = 128      # hidden dim of both encoders
D = 256      # embedding dim / decoder hidden dim
NODE_F = 7   # node feature dim   (env.py:25-28)
TOUR_F = 5   # tour feature dim


def _conv_block(prefix: str) -> Dict[str, Tuple[int, ...]]:
    return OrderedDict([
        (f"{prefix}.conv.lin_rel.weight", (H, H)),
        (f"{prefix}.conv.lin_rel.bias", (H,)),
        (f"{prefix}.conv.lin_root.weight", (H, H)),
        (f"{prefix}.norm.weight", (H,)),
        (f"{prefix}.norm.bias", (H,)),
    ])


def state_dict_schema() -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["node_encoder.input_proj.weight"] = (H, NODE_F)
    s["node_encoder.input_proj.bias"] = (H,)
    for i in range(3):
        s.update(_conv_block(f"node_encoder.layers.{i}"))
    s["node_encoder.output_proj.weight"] = (D, H)
    s["node_encoder.output_proj.bias"] = (D,)
    s["tour_encoder.input_proj.weight"] = (H, TOUR_F)
    s["tour_encoder.input_proj.bias"] = (H,)
    s["tour_encoder.node_emb_input_proj.weight"] = (H, D)
    s["tour_encoder.node_emb_input_proj.bias"] = (H,)
    for i in range(3):
        s.update(_conv_block(f"tour_encoder.tour_layers.{i}"))
    for i in range(3):
        s.update(_conv_block(f"tour_encoder.rev_tour_layers.{i}"))
    s["tour_encoder.output_proj.weight"] = (D, 2 * H)
    s["tour_encoder.output_proj.bias"] = (D,)
    s["tour_encoder.node_emb_output_proj.weight"] = (D, H)
    s["tour_encoder.node_emb_output_proj.bias"] = (D,)
    s["decoder.ctxt_proj.weight"] = (D, 2 * D)
    s["decoder.set_proj.weight"] = (3 * D, D)
    s["decoder.out_proj.weight"] = (D, D)
    return s


# older PyG releases named the two GraphConv linears lin_l / lin_r
_ALIASES = ((".conv.lin_l.", ".conv.lin_rel."), (".conv.lin_r.", ".conv.lin_root."))


def canonical_key(key: str) -> str:
    for old, new in _ALIASES:
        key = key.replace(old, new)
    return key


def random_state_dict(seed: int = 0, out_proj_scale: float = 0.1) -> "OrderedDict[str, torch.Tensor]":
    rng = np.random.default_rng(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    schema = state_dict_schema()
    for key, shape in schema.items():
        if ".norm.weight" in key:
            w = 1.0 + 0.1 * rng.uniform(-1.0, 1.0, size=shape)
        elif ".norm.bias" in key:
            w = 0.1 * rng.uniform(-1.0, 1.0, size=shape)
        else:
            wkey = key.replace(".bias", ".weight")
            bound = 1.0 / np.sqrt(schema[wkey][1])
            w = rng.uniform(-bound, bound, size=shape)
        sd[key] = torch.from_numpy(w.astype(np.float32))
    sd["decoder.out_proj.weight"] = sd["decoder.out_proj.weight"] * np.float32(out_proj_scale)
    return sd
