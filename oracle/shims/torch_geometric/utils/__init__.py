import torch
from torch_scatter import scatter


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """PyG 2.2.0 utils.softmax: segment max-subtracted exp, denominator + 1e-16."""
    assert ptr is None and dim == 0
    index = index.long()
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    src_max = scatter(src.detach(), index, 0, dim_size=N, reduce="max")
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter(out, index, 0, dim_size=N, reduce="sum") + 1e-16
    return out / out_sum.index_select(0, index)


def degree(index, num_nodes=None, dtype=None):
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros((N,), dtype=dtype or torch.get_default_dtype())
    return out.scatter_add_(0, index, torch.ones_like(index, dtype=out.dtype))


def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    """Sort by (row, col) and drop duplicate edges (enough for calc_ppr_scores.py:110)."""
    assert edge_attr is None
    N = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    key = torch.unique(edge_index[0] * N + edge_index[1], sorted=True)
    return torch.stack([key // N, key % N])


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    """PyG 2.2.0 utils.to_undirected: both directions, sorted, duplicates merged (attributes summed)."""
    if isinstance(edge_attr, int):
        edge_attr, num_nodes = None, edge_attr
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    if edge_attr is None:
        return coalesce(both, num_nodes=num_nodes)
    assert reduce == "add"
    N = int(both.max()) + 1 if num_nodes is None else num_nodes
    uniq, inv = torch.unique(both[0] * N + both[1], sorted=True, return_inverse=True)
    attr = torch.cat([edge_attr, edge_attr], dim=0)
    out = torch.zeros((uniq.numel(),) + tuple(attr.shape[1:]), dtype=attr.dtype).index_add_(0, inv, attr)
    return torch.stack([uniq // N, uniq % N]), out
