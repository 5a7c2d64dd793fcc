import torch


def dense_to_sparse(adj):
    idx = adj.nonzero(as_tuple=False).t().contiguous()
    return idx, adj[idx[0], idx[1]]
