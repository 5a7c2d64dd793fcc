"""CPU oracle of the JAMPR policy forward pass (TEST INFRASTRUCTURE ONLY, see env_oracle.py).

Plain torch-on-CPU fp32 restatement of the reference policy for the ``gnn_attn.yaml`` architecture,
written against the oracle environment's state (per-instance static data, pred/succ tour links).
Pinned by ``tests/test_oracle_golden.py`` against log-probabilities recorded from the reference
(tolerance 2e-5 absolute on log-probs, stated in the test).

Reference lines restated (all under /root/reference/lib/model):
  encode_nodes()   encoders/node_encoder.py:128-155
  _block()         encoders/graph_conv.py:66-90 around PyG GraphConv(aggr="max") — third-party,
                   pyg 2.0.3 (requirements.yml:113), absent from the mount; published semantics:
                   out_i = lin_rel(max_{j->i} w_ji x_j) + lin_root(x_i), no in-edge -> 0
  tour_encode()    encoders/tour_encoder.py:159-220, per-tour mean :226-243
  decode()         policy.py:170-201 (query / action embeddings), decoders/attn_decoder.py:62-98
  select()         policy.py:203-227
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

NEG_INF = float("-inf")


class OraclePolicy:
    def __init__(self, state_dict: Dict[str, torch.Tensor], num_heads: int = 4, clip: float = 10.0):
        self.w = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items()}
        self.heads = num_heads
        self.clip = clip
        self.reset_static()

    def reset_static(self):
        self.static_emb = None      # (I, N, 256) node-encoder output
        self.h_cache = None         # (bs, N, 128) tour-graph node embedding of the last refresh

    # ---- GraphConv block: conv -> GELU(erf) -> +skip -> LayerNorm
    def _block(self, prefix: str, x: torch.Tensor, agg: torch.Tensor) -> torch.Tensor:
        w = self.w
        y = F.linear(agg, w[f"{prefix}.conv.lin_rel.weight"], w[f"{prefix}.conv.lin_rel.bias"]) \
            + F.linear(x, w[f"{prefix}.conv.lin_root.weight"])
        y = F.gelu(y) + x
        return F.layer_norm(y, (y.size(-1),), w[f"{prefix}.norm.weight"], w[f"{prefix}.norm.bias"])

    # ---- max aggregation over the static neighbourhood graph (graph_utils.py:127-141)
    @staticmethod
    def _agg_nbh(env, x: torch.Tensor, inst: torch.Tensor) -> torch.Tensor:
        B, N, H = x.shape
        k = env.k
        knn = env.knn[inst][:, 1:, :]                                   # (B, N-1, k) sources of customer i
        wgt = env.dist[inst][:, 1:, :].gather(2, knn)                    # dist(i, source)
        src = x[torch.arange(B)[:, None, None], knn]                     # (B, N-1, k, H)
        cust = (wgt[..., None] * src).max(dim=2).values
        depot = (env.ttd[inst][:, :, None] * x).max(dim=1, keepdim=True).values   # every node -> depot
        return torch.cat((depot, cust), dim=1)

    # ---- max aggregation over the tour graph, forward (pred -> node) or reversed
    @staticmethod
    def _agg_tour(env, x: torch.Tensor, reverse: bool) -> torch.Tensor:
        B, N, H = x.shape
        ins = env.inst
        bidx = torch.arange(B)[:, None]
        nodes = torch.arange(N)[None, :].expand(B, N)
        vis = env.visited
        link = env.succ if reverse else env.pred
        w = env.dist[ins[:, None], nodes, link]
        agg = torch.where(vis[..., None], w[..., None] * x[bidx, link], torch.zeros(()))
        # depot: forward = last node of every tour, reversed = first node of every tour
        ends = vis & ((env.pred if reverse else env.succ) == 0)
        cand = torch.where(ends[..., None], env.ttd[ins][..., None] * x, torch.full((), NEG_INF))
        dep = cand.max(dim=1).values
        dep = torch.where(ends.any(-1, keepdim=True), dep, torch.zeros(()))
        agg[:, 0, :] = dep
        return agg

    def encode_nodes(self, env) -> torch.Tensor:
        w = self.w
        inst = torch.arange(env.I)
        x = F.linear(env.node_features(), w["node_encoder.input_proj.weight"], w["node_encoder.input_proj.bias"])
        for i in range(3):
            x = self._block(f"node_encoder.layers.{i}", x, self._agg_nbh(env, x, inst))
        self.static_emb = F.linear(x, w["node_encoder.output_proj.weight"], w["node_encoder.output_proj.bias"])
        return self.static_emb

    def tour_encode(self, env):
        w = self.w
        if self.static_emb is None:
            self.encode_nodes(env)
        emb = self.static_emb[env.inst]                                   # (bs, N, 256)
        if env.graph_fresh or self.h_cache is None:
            h = F.linear(emb, w["tour_encoder.node_emb_input_proj.weight"], w["tour_encoder.node_emb_input_proj.bias"])
            for i in range(3):
                agg = self._agg_nbh(env, h, env.inst) if i == 2 else self._agg_tour(env, h, False)
                h = self._block(f"tour_encoder.tour_layers.{i}", h, agg)
            for i in range(3):
                agg = self._agg_nbh(env, h, env.inst) if i == 2 else self._agg_tour(env, h, True)
                h = self._block(f"tour_encoder.rev_tour_layers.{i}", h, agg)
            self.h_cache = h
        h = self.h_cache
        # per-tour embedding: mean over the buffer positions up to and including the first depot entry
        rows = env.active_rows()                                          # (bs, K, L)
        bs, K, L = rows.shape
        idx = (rows > 0).to(torch.int).argmin(dim=-1)
        gathered = h[torch.arange(bs)[:, None, None], rows]               # (bs, K, L, H)
        csum = torch.cumsum(gathered, dim=2)
        per_tour = csum.gather(2, idx[:, :, None, None].expand(bs, K, 1, h.size(-1))).squeeze(2) / (idx + 1)[:, :, None]
        tf = F.linear(env.tour_features(), w["tour_encoder.input_proj.weight"], w["tour_encoder.input_proj.bias"])
        tour_emb = F.linear(torch.cat((tf, per_tour), -1), w["tour_encoder.output_proj.weight"],
                            w["tour_encoder.output_proj.bias"])
        node_emb = F.linear(h, w["tour_encoder.node_emb_output_proj.weight"],
                            w["tour_encoder.node_emb_output_proj.bias"]) + emb
        return tour_emb, node_emb

    def decode(self, tour_emb, node_emb, nbh, mask) -> torch.Tensor:
        """log-probabilities (bs, K*k) over (vehicle slot, candidate) pairs."""
        w = self.w
        bs, K, k = nbh.shape
        D = node_emb.size(-1)
        Hh = self.heads
        query = torch.cat((node_emb.mean(1) + tour_emb.mean(1),
                           node_emb.max(1).values + tour_emb.max(1).values), -1)
        act = (tour_emb[:, :, None, :] + node_emb[torch.arange(bs)[:, None, None], nbh]).view(bs, K * k, D)
        m = mask.reshape(bs, K * k)
        ctxt = F.linear(query, w["decoder.ctxt_proj.weight"]).view(bs, Hh, D // Hh)
        gk, gv, lk = F.linear(act, w["decoder.set_proj.weight"]).chunk(3, dim=-1)
        gk = gk.view(bs, K * k, Hh, D // Hh)
        gv = gv.view(bs, K * k, Hh, D // Hh)
        s = torch.einsum("bhd,bahd->bha", ctxt, gk) * (float(D // Hh) ** -0.5)
        s = s.masked_fill(m[:, None, :], NEG_INF)
        g = torch.einsum("bha,bahd->bhd", F.softmax(s, dim=-1), gv).reshape(bs, D)
        x = F.linear(g, w["decoder.out_proj.weight"])
        logits = torch.einsum("bd,bad->ba", x, lk) * (float(D) ** -0.5)
        logits = torch.tanh(logits) * self.clip
        logits = logits.masked_fill(m, NEG_INF)
        return F.log_softmax(logits, dim=-1)

    def forward(self, env, nbh, mask):
        tour_emb, node_emb = self.tour_encode(env)
        return self.decode(tour_emb, node_emb, nbh, mask)


def select_greedy(logp: torch.Tensor, nbh: torch.Tensor):
    """First-index argmax; action = (a // k, nbh.view(bs, A)[a]) (policy.py:211,221-223)."""
    bs, K, k = nbh.shape
    a = logp.argmax(-1)
    node = nbh.reshape(bs, K * k)[torch.arange(bs), a]
    return torch.stack((a // k, node), -1), logp.gather(-1, a[:, None])


def select_inverse_cdf(logp: torch.Tensor, nbh: torch.Tensor, u: torch.Tensor):
    """Sampling by inverse CDF on exp(logp) with given uniforms u (bs,) in [0, 1): the first index whose
    running probability sum exceeds u (last feasible index if rounding leaves the sum below u).  The
    reference samples with torch.multinomial (policy.py:213-215), whose stream cannot be reproduced
    across devices; distributional parity only."""
    bs, K, k = nbh.shape
    p = logp.exp()
    c = torch.cumsum(p, -1)
    hit = (c > u[:, None]) & (p > 0)
    a = torch.where(hit.any(-1), hit.to(torch.int).argmax(-1),
                    (p > 0).to(torch.int).cumsum(-1).argmax(-1))
    node = nbh.reshape(bs, K * k)[torch.arange(bs), a]
    return torch.stack((a // k, node), -1), logp.gather(-1, a[:, None])
