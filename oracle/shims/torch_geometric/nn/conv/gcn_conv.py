import torch
from torch.nn import Parameter
from ..dense.linear import Linear
from .. import inits


def gcn_norm_sparse(adj_t):
    """PyG 2.2.0 gcn_norm for a SparseTensor adj_t with add_self_loops=True:
    missing values -> 1, diagonal *set* to 1 (fill_diag), D^-1/2 A D^-1/2."""
    adj = adj_t.coalesce() if adj_t.layout == torch.sparse_coo else adj_t.to_sparse_coo().coalesce()
    N = adj.size(0)
    row, col = adj.indices()
    val = adj.values().to(torch.float32)
    keep = row != col
    ar = torch.arange(N)
    row = torch.cat([row[keep], ar])
    col = torch.cat([col[keep], ar])
    val = torch.cat([val[keep], torch.ones(N)])
    A = torch.sparse_coo_tensor(torch.stack([row, col]), val, (N, N)).coalesce()
    row, col = A.indices()
    val = A.values()
    deg = torch.zeros(N).scatter_add_(0, row, val)
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0.0
    val = (val * dis[row]) * dis[col]
    return torch.sparse_coo_tensor(torch.stack([row, col]), val, (N, N)).coalesce()


class GCNConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 add_self_loops=True, normalize=True, bias=True, **kwargs):
        super().__init__()
        assert normalize and add_self_loops and not improved
        self.cached = cached
        self._cached_adj_t = None
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        self.bias = Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        self.lin.reset_parameters()
        inits.zeros(self.bias)
        self._cached_adj_t = None

    def forward(self, x, adj_t, edge_weight=None):
        cache = self._cached_adj_t
        if cache is None:
            A = gcn_norm_sparse(adj_t)
            if self.cached:
                self._cached_adj_t = A
        else:
            A = cache
        x = self.lin(x)
        out = torch.sparse.mm(A, x)
        return out + self.bias
