"""CPU emulation of the tensor-core policy step's roundings (no GPU needed): which of the 16-bit roundings of
csrc/policy_tc2.cu cost how much log-probability accuracy.  Teacher-forced on a golden trace of the reference.

Every GEMM keeps fp32 accumulation (as tcgen05 does); what is switchable is where values are rounded to the 16-bit
operand type or split into hi + lo parts.  Switches (comma separated in argv[3], default = the round-1 kernel):
  act16    GraphConv A operands (h, agg) rounded                      [on]
  w16      GraphConv / projection B operands (weights) rounded        [on]
  agg16    max-aggregation computed on 16-bit h x 16-bit edge weight, product rounded   [on]
  ne16     node embedding parked in 16 bits for the decoder           [on]
  dec16    decoder mat-vec weights in 16 bits                         [on]
  enc16    node encoder (static_emb, h0) through the same 16-bit GEMMs [on]
  actsplit A operands as hi + lo (two 16-bit terms)                   [off]
  wsplit   B operands as hi + lo                                      [off]
Usage: python tools/emulate_tc_precision.py g101_pomo fp16 act16,w16,agg16,ne16,dec16,enc16 [max_steps]"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.policy_oracle import NEG_INF, OraclePolicy  # noqa: E402
from oracle.replay import load_case, make_oracle, start_nodes_of  # noqa: E402


class EmuPolicy(OraclePolicy):
    def __init__(self, sd, dtype, sw):
        super().__init__(sd)
        self.dt = dtype
        self.sw = set(sw)

    def r(self, x):
        return x.to(self.dt).to(torch.float32)

    def opA(self, x):
        if "actsplit" in self.sw:
            hi = self.r(x)
            return hi + self.r(x - hi)
        return self.r(x) if "act16" in self.sw else x

    def opB(self, w):
        if "wsplit" in self.sw:
            hi = self.r(w)
            return hi + self.r(w - hi)
        return self.r(w) if "w16" in self.sw else w

    def _block(self, prefix, x, agg):
        w = self.w
        y = F.linear(self.opA(agg), self.opB(w[f"{prefix}.conv.lin_rel.weight"]), w[f"{prefix}.conv.lin_rel.bias"]) \
            + F.linear(self.opA(x), self.opB(w[f"{prefix}.conv.lin_root.weight"]))
        y = F.gelu(y) + x
        return F.layer_norm(y, (y.size(-1),), w[f"{prefix}.norm.weight"], w[f"{prefix}.norm.bias"])

    def _agg_nbh_e(self, env, x, inst):
        if "agg16" not in self.sw:
            return OraclePolicy._agg_nbh(env, self.opA(x) if "act16" in self.sw or "actsplit" in self.sw else x, inst)
        B, N, H = x.shape
        xr = self.r(x)
        knn = env.knn[inst][:, 1:, :]
        wgt = self.r(env.dist[inst][:, 1:, :].gather(2, knn))
        src = xr[torch.arange(B)[:, None, None], knn]
        cust = self.r(wgt[..., None] * src).max(dim=2).values
        depot = self.r(self.r(env.ttd[inst])[:, :, None] * xr).max(dim=1, keepdim=True).values
        return torch.cat((depot, cust), dim=1)

    def _agg_tour_e(self, env, x, reverse):
        if "agg16" not in self.sw:
            return OraclePolicy._agg_tour(env, self.opA(x) if "act16" in self.sw or "actsplit" in self.sw else x, reverse)
        B, N, H = x.shape
        xr = self.r(x)
        ins = env.inst
        bidx = torch.arange(B)[:, None]
        nodes = torch.arange(N)[None, :].expand(B, N)
        vis = env.visited
        link = env.succ if reverse else env.pred
        w = self.r(env.dist[ins[:, None], nodes, link])
        agg = torch.where(vis[..., None], self.r(w[..., None] * xr[bidx, link]), torch.zeros(()))
        ends = vis & ((env.pred if reverse else env.succ) == 0)
        cand = torch.where(ends[..., None], self.r(self.r(env.ttd[ins])[..., None] * xr), torch.full((), NEG_INF))
        dep = cand.max(dim=1).values
        dep = torch.where(ends.any(-1, keepdim=True), dep, torch.zeros(()))
        agg[:, 0, :] = dep
        return agg

    def encode_nodes(self, env):
        if "enc16" not in self.sw:
            return super().encode_nodes(env)
        w = self.w
        inst = torch.arange(env.I)
        x = F.linear(env.node_features(), w["node_encoder.input_proj.weight"], w["node_encoder.input_proj.bias"])
        for i in range(3):
            x = self._block(f"node_encoder.layers.{i}", x, self._agg_nbh_e(env, x, inst))
        self.static_emb = F.linear(self.opA(x), self.opB(w["node_encoder.output_proj.weight"]),
                                   w["node_encoder.output_proj.bias"])
        return self.static_emb

    def forward(self, env, nbh, mask):
        w = self.w
        if self.static_emb is None:
            self.encode_nodes(env)
        emb = self.static_emb[env.inst]
        if env.graph_fresh or self.h_cache is None:
            if "enc16" in self.sw:
                h = F.linear(self.opA(emb), self.opB(w["tour_encoder.node_emb_input_proj.weight"]),
                             w["tour_encoder.node_emb_input_proj.bias"])
            else:
                h = F.linear(emb, w["tour_encoder.node_emb_input_proj.weight"], w["tour_encoder.node_emb_input_proj.bias"])
            for i in range(3):
                agg = self._agg_nbh_e(env, h, env.inst) if i == 2 else self._agg_tour_e(env, h, False)
                h = self._block(f"tour_encoder.tour_layers.{i}", h, agg)
            for i in range(3):
                agg = self._agg_nbh_e(env, h, env.inst) if i == 2 else self._agg_tour_e(env, h, True)
                h = self._block(f"tour_encoder.rev_tour_layers.{i}", h, agg)
            self.h_cache = h
        h = self.h_cache
        hq = self.opA(h)                                   # what sits in the operand tiles
        rows = env.active_rows()
        bs, K, L = rows.shape
        idx = (rows > 0).to(torch.int).argmin(dim=-1)
        gathered = hq[torch.arange(bs)[:, None, None], rows]
        csum = torch.cumsum(gathered, dim=2)
        per_tour = csum.gather(2, idx[:, :, None, None].expand(bs, K, 1, h.size(-1))).squeeze(2) / (idx + 1)[:, :, None]
        def dwn(tag):
            return (lambda m: self.r(m)) if ("dec16" in self.sw or tag in self.sw) else (lambda m: m)
        wto = w["tour_encoder.output_proj.weight"].double()
        tof_w = (wto[:, :128] @ w["tour_encoder.input_proj.weight"].double()).float()
        tof_b = (wto[:, :128] @ w["tour_encoder.input_proj.bias"].double() + w["tour_encoder.output_proj.bias"].double()).float()
        te = F.linear(env.tour_features(), tof_w, tof_b) + F.linear(per_tour, dwn("d_to2")(w["tour_encoder.output_proj.weight"][:, 128:]))
        ne = F.linear(hq, self.opB(w["tour_encoder.node_emb_output_proj.weight"]),
                      w["tour_encoder.node_emb_output_proj.bias"]) + emb
        self.last_h, self.last_ne = h, ne
        nep = self.r(ne) if "ne16" in self.sw else ne
        # decoder in the kernel's algebra
        k = nbh.size(-1)
        A = K * k
        m = mask.reshape(bs, A)
        q = torch.cat((nep.mean(1) + te.mean(1), nep.max(1).values + te.max(1).values), -1)
        ctxt = F.linear(q, dwn("d_ctxt")(w["decoder.ctxt_proj.weight"]))                  # (bs, 256)
        sp = w["decoder.set_proj.weight"]
        gk, gv, lk = sp[0:256], sp[256:512], sp[512:768]
        # qp[h] = W_gk,h^T ctxt_h
        qp = torch.einsum("bhd,hdc->bhc", ctxt.view(bs, 4, 64), dwn("d_gk")(gk).view(4, 64, 256))
        act_n = nep[torch.arange(bs)[:, None, None], nbh].view(bs, A, 256)
        act_t = te[:, :, None, :].expand(bs, K, k, 256).reshape(bs, A, 256)
        s = (torch.einsum("bhc,bac->bha", qp, act_n) + torch.einsum("bhc,bac->bha", qp, act_t)) * 0.125
        s = s.masked_fill(m[:, None, :], NEG_INF)
        p = F.softmax(s, dim=-1)
        gbar = torch.einsum("bha,bac->bhc", p, act_n + act_t)                  # (bs, 4, 256)
        g = torch.einsum("hdc,bhc->bhd", dwn("d_gv")(gv).view(4, 64, 256), gbar).reshape(bs, 256)
        m1 = (lk.double().t() @ w["decoder.out_proj.weight"].double()).float()  # lp = W_lk^T W_out g
        lp = F.linear(g, dwn("d_m1")(m1))
        z = torch.einsum("bc,bac->ba", lp, act_n) + torch.einsum("bc,bac->ba", lp, act_t)
        self.last_lp = lp
        logits = (torch.tanh(z * 0.0625) * 10.0).masked_fill(m, NEG_INF)
        return F.log_softmax(logits, dim=-1)


def run(name, dtype, sw, max_steps=None):
    g = load_case(name)
    env, ref_pol = make_oracle(g)
    pol = EmuPolicy(ref_pol.w, dtype, sw)
    env.reset(start_nodes=start_nodes_of(g))
    T = g["trace_action"].shape[0] if max_steps is None else min(max_steps, g["trace_action"].shape[0])
    errs, rels, gaps, agree, herr, neerr, lpn = [], [], [], [], [], [], []
    for t in range(T):
        nbh, mask = env.observe()
        logp = pol.forward(env, nbh, mask)
        ref_lp = ref_pol.forward(env, nbh, mask)
        te_r, ne_r = ref_pol.tour_encode(env)
        herr.append(float((pol.last_h - ref_pol.h_cache).abs().max()))
        neerr.append(float((pol.last_ne - ne_r).abs().max()))
        lpn.append(float(pol.last_lp.norm(dim=-1).max()))
        ref = g["trace_logp"][t]
        fin = np.isfinite(ref)
        lp = logp.numpy()
        d = np.abs(lp - ref)[fin]
        errs.append(d.max())
        rels.append((d / np.maximum(np.abs(ref[fin]), 1.0)).max())
        srt = np.sort(ref, -1)
        gaps.append(srt[:, -1] - srt[:, -2])
        a = lp.argmax(-1)
        aref = ref.argmax(-1)
        agree.append(a == aref)
        env.step(torch.from_numpy(g["trace_action"][t].astype(np.int64)))
    errs, gaps, agree = np.array(errs), np.concatenate(gaps), np.concatenate(agree)
    out = {"case": name, "dtype": str(dtype), "switches": sorted(sw), "steps": len(errs),
           "logp_abs_max": float(errs.max()), "logp_abs_med_of_step_max": float(np.median(errs)),
           "logp_rel_max(|ref|>=1)": float(np.max(rels)),
           "h_abs_max": float(np.max(herr)), "ne_abs_max": float(np.max(neerr)), "lp_norm_max": float(np.max(lpn)),
           "flips": int((~agree).sum()), "flips_gap>1e-3": int((~agree[gaps > 1e-3]).sum()), "decisions": int(agree.size)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "g101_pomo"
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[2] if len(sys.argv) > 2 else "fp16"]
    sws = sys.argv[3] if len(sys.argv) > 3 else "act16,w16,agg16,ne16,dec16,enc16"
    ms = int(sys.argv[4]) if len(sys.argv) > 4 else None
    torch.set_num_threads(8)
    for s in sws.split("/"):
        run(name, dt, [x for x in s.split(",") if x and x != "none"], ms)
