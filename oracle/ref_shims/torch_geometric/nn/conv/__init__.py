"""GraphConv / MessagePassing restated from the published PyG 2.0.3 semantics."""
import torch
import torch.nn as nn


class MessagePassing(nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        self.aggr = aggr


class GraphConv(MessagePassing):
    def __init__(self, in_channels, out_channels, aggr="add", bias=True, **kwargs):
        super().__init__(aggr=aggr)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.lin_rel = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_root = nn.Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x, edge_index, edge_weight=None, size=None):
        src, dst = edge_index[0], edge_index[1]
        msg = x[src] if edge_weight is None else edge_weight.view(-1, 1) * x[src]
        out = torch.zeros_like(x)
        idx = dst.view(-1, 1).expand_as(msg)
        if self.aggr == "max":
            out = out.scatter_reduce(0, idx, msg, reduce="amax", include_self=False)
        elif self.aggr in ("add", "sum"):
            out = out.scatter_add(0, idx, msg)
        elif self.aggr == "mean":
            out = out.scatter_reduce(0, idx, msg, reduce="mean", include_self=False)
        else:
            raise ValueError(self.aggr)
        return self.lin_rel(out) + self.lin_root(x)
