import inspect
import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    """The slice of PyG 2.2.0 MessagePassing that LinkAttention uses: bipartite
    x=(x[0], x[1]) with x[k] belonging to edge_index[k], node_dim=0, aggr='add'.
    flow='target_to_source' makes (i, j) = (0, 1): `_i` tensors are lifted by
    edge_index[0], `_j` by edge_index[1]; messages are summed over edge_index[i]
    into x[i].size(0) rows."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert aggr == "add" and node_dim == 0
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        x = kwargs.pop("x")
        args = {
            "x_i": x[i].index_select(0, edge_index[i]),
            "x_j": x[j].index_select(0, edge_index[j]),
            "index": edge_index[i],
            "ptr": None,
            "size_i": x[i].size(0),
        }
        args.update(kwargs)
        params = inspect.signature(self.message).parameters
        msg = self.message(**{k: args[k] for k in params})
        return scatter(msg, edge_index[i], dim=0, dim_size=args["size_i"], reduce="sum")
