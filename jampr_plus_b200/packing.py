"""Checkpoint state_dict -> device weight blob(s) consumed by libjampr_b200.so.

The fp32 blob layout is owned by the library (csrc/common.cuh, queried through jampr_weight_offsets);
GEMM operands are stored K-major ("WT": row = input feature) so that kernels stream them row by row."""
from typing import Dict

import torch

from . import _cabi
from .weights import canonical_key, state_dict_schema


def validate_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Canonical key names (older PyG lin_l/lin_r aliases accepted), exact schema check (fails loudly)."""
    schema = state_dict_schema()
    out = {canonical_key(k): v for k, v in sd.items()}
    missing = [k for k in schema if k not in out]
    extra = [k for k in out if k not in schema]
    if missing or extra:
        raise KeyError(f"state_dict does not match the gnn_attn policy schema: missing={missing} unexpected={extra}")
    for k, shape in schema.items():
        if tuple(out[k].shape) != tuple(shape):
            raise ValueError(f"{k}: expected shape {tuple(shape)}, got {tuple(out[k].shape)}")
    return out


def _block(sd, prefix: str) -> torch.Tensor:
    wt = torch.cat((sd[f"{prefix}.conv.lin_rel.weight"].t(), sd[f"{prefix}.conv.lin_root.weight"].t()), 0)  # (256,128)
    return torch.cat((wt.reshape(-1), sd[f"{prefix}.conv.lin_rel.bias"], sd[f"{prefix}.norm.weight"],
                      sd[f"{prefix}.norm.bias"]))


def pack_f32(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """CPU fp32 blob in the layout of csrc/common.cuh."""
    sd = {k: v.detach().to(torch.float32).cpu() for k, v in validate_state_dict(sd).items()}
    off = _cabi.weight_offsets()
    blob = torch.zeros(off["END"], dtype=torch.float32)

    def put(name, t):
        t = t.contiguous().reshape(-1)
        blob[off[name]: off[name] + t.numel()] = t

    def halves(w):  # (256, K) -> [2][K][128]: column halves of W^T
        wt = w.t()  # (K, 256)
        return torch.stack((wt[:, :128], wt[:, 128:]), 0)

    w_in = torch.zeros(8, 128)
    w_in[:7] = sd["node_encoder.input_proj.weight"].t()
    put("NE_IN_WT", w_in)
    put("NE_IN_B", sd["node_encoder.input_proj.bias"])
    put("NE_BLK", torch.cat([_block(sd, f"node_encoder.layers.{i}") for i in range(3)]))
    put("NE_OUT_WT", halves(sd["node_encoder.output_proj.weight"]))
    put("NE_OUT_B", sd["node_encoder.output_proj.bias"])
    put("TE_IN_WT", sd["tour_encoder.node_emb_input_proj.weight"].t())
    put("TE_IN_B", sd["tour_encoder.node_emb_input_proj.bias"])
    put("TE_BLK", torch.cat([_block(sd, f"tour_encoder.tour_layers.{i}") for i in range(3)]
                            + [_block(sd, f"tour_encoder.rev_tour_layers.{i}") for i in range(3)]))
    put("TE_OUT_WT", halves(sd["tour_encoder.node_emb_output_proj.weight"]))
    put("TE_OUT_B", sd["tour_encoder.node_emb_output_proj.bias"])
    w_tf = torch.zeros(8, 128)
    w_tf[:5] = sd["tour_encoder.input_proj.weight"].t()
    put("TF_WT", w_tf)
    put("TF_B", sd["tour_encoder.input_proj.bias"])
    put("TO_WT", sd["tour_encoder.output_proj.weight"].t())
    put("TO_B", sd["tour_encoder.output_proj.bias"])
    put("CTXT_WT", sd["decoder.ctxt_proj.weight"].t())
    sp = sd["decoder.set_proj.weight"]
    put("GK_W", sp[0:256])
    put("GV_WT", sp[256:512].t())
    put("LK_W", sp[512:768])
    put("OUT_WT", sd["decoder.out_proj.weight"].t())
    # tensor-core path: the tour-feature branch of tour_encoder.output_proj folded into a 5 x 256 matrix
    wto = sd["tour_encoder.output_proj.weight"].double()
    tof = torch.zeros(8, 256)
    tof[:5] = (wto[:, :128] @ sd["tour_encoder.input_proj.weight"].double()).t().float()
    put("TOF_WT", tof)
    put("TOF_B", (wto[:, :128] @ sd["tour_encoder.input_proj.bias"].double()
                  + sd["tour_encoder.output_proj.bias"].double()).float())
    return blob


def sw128_image(w: torch.Tensor) -> torch.Tensor:
    """Shared-memory image of a K-major UMMA operand with the 128-byte swizzle (csrc/tc_ptx.cuh).

    ``w`` is (rows, K) of a 16-bit dtype, K a multiple of 64.  The image is K/64 tiles of [rows][64 elements]
    (128 B per row); inside a tile the 16-byte chunk c of row r sits at chunk position c ^ (r & 7).  A bulk
    copy of the image into 1024-byte aligned shared memory is directly consumable by tcgen05.mma."""
    assert w.dim() == 2 and w.element_size() == 2 and w.size(1) % 64 == 0
    rows, K = w.shape
    nkb = K // 64
    t = w.contiguous().view(torch.int16).view(rows, nkb, 8, 8)            # (r, kb, chunk, 8 elements)
    r = torch.arange(rows).view(rows, 1, 1).expand(rows, nkb, 8)
    c = torch.arange(8).view(1, 1, 8).expand(rows, nkb, 8)
    src = (c ^ (r & 7))                                                   # position p holds chunk p ^ (r & 7)
    img = torch.gather(t, 2, src.unsqueeze(-1).expand(rows, nkb, 8, 8))
    return img.permute(1, 0, 2, 3).contiguous().view(-1).view(w.dtype)


def pack_tc(sd: Dict[str, torch.Tensor], dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """16-bit operands of the tensor-core policy step (layout owned by csrc/policy_tc.cu, W16_*):

    * 7 x 4 weight stages of 16 KB, each the swizzled shared-memory image of a [128 rows][64 K] tile: for the six
      GraphConv blocks (tour_layers 0..2, rev_tour_layers 0..2) the B operand is cat([lin_rel.weight,
      lin_root.weight], dim=1) (128 x 256; K-blocks 0,1 multiply the aggregated messages, 2,3 the node itself);
      then node_emb_output_proj.weight as two 128-row halves x two K-blocks;
    * decoder mat-vec matrices stored [K][256] (input-major): per-tour-mean half of tour output_proj, ctxt_proj^T,
      set_proj[0:256] as stored, set_proj[256:512]^T, and M1^T = out_proj^T @ set_proj[512:768] (computed in fp64);
    * 20 more weight stages for the large-graph path: the node-encoder blocks, its output_proj and the hoisted
      node_emb_input_proj."""
    assert dtype in (torch.bfloat16, torch.float16)
    sd = {k: v.detach().to(torch.float32).cpu() for k, v in validate_state_dict(sd).items()}
    parts = []
    for prefix in [f"tour_encoder.tour_layers.{i}" for i in range(3)] + \
                  [f"tour_encoder.rev_tour_layers.{i}" for i in range(3)]:
        wcat = torch.cat((sd[f"{prefix}.conv.lin_rel.weight"], sd[f"{prefix}.conv.lin_root.weight"]), 1)   # (128, 256)
        parts.append(sw128_image(wcat.to(dtype)))
    wo = sd["tour_encoder.node_emb_output_proj.weight"]                                                  # (256, 128)
    for half in range(2):
        parts.append(sw128_image(wo[half * 128:(half + 1) * 128].contiguous().to(dtype)))
    sp = sd["decoder.set_proj.weight"]
    m1t = (sd["decoder.out_proj.weight"].double().t() @ sp[512:768].double()).float()
    for m in (sd["tour_encoder.output_proj.weight"][:, 128:].t(), sd["decoder.ctxt_proj.weight"].t(), sp[0:256],
              sp[256:512].t(), m1t):
        parts.append(m.contiguous().to(dtype).reshape(-1))
    # large-graph path (csrc/gnn_large.cu): node-encoder GraphConv blocks, node_encoder.output_proj (two 128-row
    # halves) and tour_encoder.node_emb_input_proj (128 x 256) as 4 + 4 + 4 + 4 + 4 more weight stages
    for i in range(3):
        prefix = f"node_encoder.layers.{i}"
        wcat = torch.cat((sd[f"{prefix}.conv.lin_rel.weight"], sd[f"{prefix}.conv.lin_root.weight"]), 1)
        parts.append(sw128_image(wcat.to(dtype)))
    wno = sd["node_encoder.output_proj.weight"]
    for half in range(2):
        parts.append(sw128_image(wno[half * 128:(half + 1) * 128].contiguous().to(dtype)))
    parts.append(sw128_image(sd["tour_encoder.node_emb_input_proj.weight"].contiguous().to(dtype)))
    # split decoder (csrc/decoder_tc.cu, WD_*): B operands [rows = outputs][K] as K-block tiles, a hi part followed by a
    # lo part with w = hi + lo (both 16-bit): every dense decoder stage runs hi*hi + lo*hi + hi*lo on the tensor cores
    m1 = (sp[512:768].double().t() @ sd["decoder.out_proj.weight"].double()).float()      # lp = M1 g, M1 = W_lk^T W_out
    mats = [sd["decoder.ctxt_proj.weight"]]                                                # (256, 512): 8 tiles
    mats += [sp[64 * h:64 * h + 64, :].t().contiguous() for h in range(4)]                 # (256, 64) per head: qp_h = B ctxt_h
    mats += [sp[256 + 64 * h:256 + 64 * h + 64, :].contiguous() for h in range(4)]         # (64, 256) per head: 4 tiles each
    mats += [m1]                                                                           # (256, 256): 4 tiles
    # tour_emb[t] = W_to [input_proj(features_t) | mean_h[t]] + b as ONE product over K = 192: the per-tour-mean half
    # (128), then the folded feature half: 5 feature columns, the folded bias (multiplied by a constant-1 column), zeros
    wto64 = sd["tour_encoder.output_proj.weight"].double()
    feat = torch.zeros(256, 64, dtype=torch.float64)
    feat[:, :5] = wto64[:, :128] @ sd["tour_encoder.input_proj.weight"].double()
    feat[:, 5] = wto64[:, :128] @ sd["tour_encoder.input_proj.bias"].double() + sd["tour_encoder.output_proj.bias"].double()
    mats += [torch.cat((sd["tour_encoder.output_proj.weight"][:, 128:], feat.float()), 1).contiguous()]   # (256, 192): 3 tiles
    hi = [m.to(dtype) for m in mats]
    lo = [(m - h_.float()).to(dtype) for m, h_ in zip(mats, hi)]
    for part in (hi, lo):
        parts.extend(sw128_image(m.contiguous()) for m in part)
    blob = torch.cat(parts)
    lib = _cabi.load_library()
    assert blob.numel() == lib.jampr_weight_blob16_elems(), (blob.numel(), lib.jampr_weight_blob16_elems())
    return blob
