"""Shim of torch_sparse (TEST INFRASTRUCTURE ONLY).  The model code only forwards `adj_t` to GCNConv, where a torch
sparse tensor stands in; the reference's data-ingestion functions (util/read_datasets.py, util/calc_ppr_scores.py),
which tests/test_datasets.py runs unmodified, need a little more of `SparseTensor`: the published semantics of
torch_sparse/tensor.py `from_edge_index`, `coalesce`, `to_symmetric`, `to_torch_sparse_coo_tensor`, restated."""
import torch


class SparseTensor:
    def __init__(self, row, col, value=None, sparse_sizes=None, is_sorted=False):
        n = tuple(int(s) for s in sparse_sizes) if sparse_sizes is not None else (int(row.max()) + 1, int(col.max()) + 1)
        if not is_sorted:                      # SparseStorage sorts by (row, col), stable; duplicates are KEPT
            order = torch.argsort(row * n[1] + col, stable=True)
            row, col = row[order], col[order]
            value = None if value is None else value[order]
        self._row, self._col, self._value, self._sizes = row, col, value, n

    @classmethod
    def from_edge_index(cls, edge_index, edge_attr=None, sparse_sizes=None, is_sorted=False, trust_data=False):
        return cls(edge_index[0], edge_index[1], edge_attr, sparse_sizes, is_sorted)

    def coo(self):
        return self._row, self._col, self._value

    def sparse_sizes(self):
        return self._sizes

    def size(self, dim):
        return self._sizes[dim]

    def nnz(self):
        return self._col.numel()

    def to(self, *a, **k):
        v = None if self._value is None else self._value.to(*a, **k)
        dev = [x for x in a if isinstance(x, (str, torch.device))]
        row, col = (self._row.to(dev[0]), self._col.to(dev[0])) if dev else (self._row, self._col)
        return SparseTensor(row, col, v, self._sizes, is_sorted=True)

    def coalesce(self, reduce="sum"):
        assert reduce in ("sum", "add")
        key = self._row * self._sizes[1] + self._col
        uniq, inv = torch.unique(key, return_inverse=True)
        v = None
        if self._value is not None:
            v = torch.zeros(uniq.numel(), dtype=self._value.dtype).index_add_(0, inv, self._value)
        return SparseTensor(uniq // self._sizes[1], uniq % self._sizes[1], v, self._sizes, is_sorted=True)

    def to_symmetric(self, reduce="sum"):
        # both directions of every entry, duplicates merged by `reduce` (torch_sparse/tensor.py to_symmetric)
        n = max(self._sizes)
        row = torch.cat([self._row, self._col])
        col = torch.cat([self._col, self._row])
        v = None if self._value is None else torch.cat([self._value, self._value])
        return SparseTensor(row, col, v, (n, n)).coalesce(reduce)

    def to_torch_sparse_coo_tensor(self, dtype=None):
        v = self._value if self._value is not None else torch.ones(self.nnz(), dtype=dtype or torch.float)
        return torch.sparse_coo_tensor(torch.stack([self._row, self._col]), v, self._sizes)
