import torch


def scatter_sum(src, index, dim=0, out=None, dim_size=None):
    size = list(src.size())
    size[dim] = int(dim_size if dim_size is not None else index.max() + 1)
    res = torch.zeros(size, dtype=src.dtype, device=src.device) if out is None else out
    return res.index_add_(dim, index, src)
