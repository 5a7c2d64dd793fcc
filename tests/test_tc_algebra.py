"""Host-side check of the algebra the tensor-core policy step relies on (csrc/policy_tc.cu): the decoder restated
with the set_proj thirds moved onto the query side, the glimpse weights folded onto nodes / tours, M1 = W_lk^T W_out
and the folded tour-feature matrix must reproduce the oracle decoder (attn_decoder.py:62-98) in float64."""
import numpy as np
import torch

from jampr_plus_b200.weights import random_state_dict
from oracle.policy_oracle import OraclePolicy


def test_restructured_decoder_equals_oracle():
    torch.manual_seed(0)
    sd = random_state_dict(0)
    pol = OraclePolicy(sd)
    bs, N, K, k = 5, 23, 3, 6
    A = K * k
    ne = torch.randn(bs, N, 256)
    mean_h = torch.randn(bs, K, 128)
    feat = torch.rand(bs, K, 5)
    nbh = torch.randint(0, N, (bs, K, k))
    mask = torch.rand(bs, K, k) < 0.3
    mask[:, 0, 0] = False
    w = {k_: v.double() for k_, v in sd.items()}
    # oracle: tour_emb through the unfolded layers, then decode()
    tf = torch.nn.functional.linear(feat, sd["tour_encoder.input_proj.weight"], sd["tour_encoder.input_proj.bias"])
    te = torch.nn.functional.linear(torch.cat((tf, mean_h), -1), sd["tour_encoder.output_proj.weight"],
                                    sd["tour_encoder.output_proj.bias"])
    ref = pol.decode(te, ne, nbh, mask).double()
    # restructured, float64
    ne64, mh, ft = ne.double(), mean_h.double(), feat.double()
    wto = w["tour_encoder.output_proj.weight"]
    tofw = wto[:, :128] @ w["tour_encoder.input_proj.weight"]                 # (256, 5)
    tofb = wto[:, :128] @ w["tour_encoder.input_proj.bias"] + w["tour_encoder.output_proj.bias"]
    te2 = ft @ tofw.t() + tofb + mh @ wto[:, 128:].t()
    q = torch.cat((ne64.mean(1) + te2.mean(1), ne64.max(1).values + te2.max(1).values), -1)
    ctxt = q @ w["decoder.ctxt_proj.weight"].t()
    sp = w["decoder.set_proj.weight"]
    gk, gv, lk = sp[:256], sp[256:512], sp[512:]
    qp = torch.stack([ctxt[:, h * 64:(h + 1) * 64] @ gk[h * 64:(h + 1) * 64] for h in range(4)], 1)   # (bs,4,256)
    sn = torch.einsum("bhc,bnc->bhn", qp, ne64)
    qt = torch.einsum("bhc,btc->bht", qp, te2)
    flat = nbh.reshape(bs, A)
    tour_of = torch.arange(A) // k
    sc = (sn.gather(2, flat[:, None, :].expand(bs, 4, A)) + qt[:, :, tour_of]) * 0.125
    sc = sc.masked_fill(mask.reshape(bs, 1, A), float("-inf"))
    p = torch.softmax(sc, -1)
    pn = torch.zeros(bs, 4, N, dtype=torch.float64).scatter_add_(2, flat[:, None, :].expand(bs, 4, A), p)
    P = p.view(bs, 4, K, k).sum(-1)
    gbar = torch.einsum("bht,btc->bhc", P, te2) + torch.einsum("bhn,bnc->bhc", pn, ne64)
    g = torch.cat([gbar[:, h] @ gv[h * 64:(h + 1) * 64].t() for h in range(4)], -1)
    m1t = w["decoder.out_proj.weight"].t() @ lk
    lp = g @ m1t
    logit = 10.0 * torch.tanh((torch.einsum("bc,bnc->bn", lp, ne64).gather(1, flat)
                               + torch.einsum("bc,btc->bt", lp, te2)[:, tour_of]) / 16.0)
    logit = logit.masked_fill(mask.reshape(bs, A), float("-inf"))
    mine = torch.log_softmax(logit, -1)
    fin = torch.isfinite(ref)
    assert torch.equal(fin, torch.isfinite(mine))
    assert (mine[fin] - ref[fin]).abs().max().item() < 2e-5      # fp32 oracle vs fp64 restatement


def test_pack_tc_layout():
    from jampr_plus_b200.packing import pack_tc, sw128_image
    sd = random_state_dict(0)
    blob = pack_tc(sd, torch.bfloat16)
    # first stage = swizzled image of lin_rel.weight[:, 0:64] of tour_layers.0 (K-block 0 of [lin_rel | lin_root])
    w = sd["tour_encoder.tour_layers.0.conv.lin_rel.weight"][:, :64].contiguous().to(torch.bfloat16)
    assert torch.equal(blob[:8192].view(torch.int16), sw128_image(w).view(torch.int16))
    # stage 2 of the same layer = lin_root.weight[:, 0:64]
    w = sd["tour_encoder.tour_layers.0.conv.lin_root.weight"][:, :64].contiguous().to(torch.bfloat16)
    assert torch.equal(blob[2 * 8192:3 * 8192].view(torch.int16), sw128_image(w).view(torch.int16))
    # layer 6, stage 3 = node_emb_output_proj.weight[128:256, 64:128]
    w = sd["tour_encoder.node_emb_output_proj.weight"][128:, 64:].contiguous().to(torch.bfloat16)
    assert torch.equal(blob[27 * 8192:28 * 8192].view(torch.int16), sw128_image(w).view(torch.int16))
