"""Shim of torch_scatter.scatter (test infrastructure; see oracle/shims/README.md)."""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0
    index = index.long()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = (dim_size,) + tuple(src.shape[1:])
    if reduce in ("sum", "add"):
        res = torch.zeros(shape, dtype=src.dtype, device=src.device)
        idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
        return res.scatter_add_(0, idx, src)
    if reduce == "max":
        res = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
        idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
        return res.scatter_reduce_(0, idx, src, reduce="amax", include_self=True)
    raise NotImplementedError(reduce)
