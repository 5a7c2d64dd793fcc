"""knn_graph restated (see ../README.md for the fixed tie-break)."""
import torch


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", cosine=False, num_workers=1):
    assert batch is None and not cosine
    n = x.size(0)
    d = x[:, None, :] - x[None, :, :]
    d2 = (d * d).sum(-1)           # for 2-D points: dx*dx + dy*dy, one rounding per op
    if not loop:
        d2 = d2 + torch.diag(torch.full((n,), float("inf"), dtype=d2.dtype))
    order = torch.sort(d2, dim=-1, stable=True).indices[:, :k]   # ties -> lower index
    centre = torch.arange(n, device=x.device)[:, None].expand(n, k)
    return torch.stack((order.reshape(-1), centre.reshape(-1)), dim=0)
