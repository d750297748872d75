"""Eval-side drivers around LinkTransformer: link-sharded multi-GPU scoring.

Links are independent units, so candidate links are sharded across ranks with NO per-batch
collective.  The only exchange is per eval: the last GCN layer (+ gnn_norm + the K/V
projection) is computed for a row shard on each rank and the node tables [X | KV] are
replicated by ONE all-gather over NCCL/NVLink.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


@torch.no_grad()
def propagate_replicated(model, test_set=False, group=None):
    """X_node [N, dim] on every rank, with the K/V tables of all attention layers primed.

    world_size 1: plain model.propagate().  Otherwise GCN layers 0..L-2 are computed redundantly,
    the last layer's SpMM + LN/ReLU/residual + gnn_norm + K/V projection only for this rank's rows,
    followed by a single all_gather_into_tensor of the packed [X | KV_0 | ...] rows."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        X = model.propagate(test_set=test_set)
        model._get_kv(X)
        return X
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = model._dev()
    adj = model.get_adj(test_set)
    n, d = adj.n, model.dim
    x = model.data["x"]
    if "emb" in model.data:
        x = model.data["emb"](x)
    x = x.detach().to(dev, torch.float32)
    gcn = model.node_encoder.gnn_encoder
    convs = list(gcn.convs)
    per = (n + world - 1) // world
    r0 = min(n, rank * per)
    rows = min(n, r0 + per) - r0
    for i, conv in enumerate(convs):
        last = i == len(convs) - 1
        ln = gcn.lns[i] if gcn.lns is not None else None
        res_ok = gcn.residual and x.shape[-1] == conv.bias.numel()
        if not last:
            xi = conv(x, adj)
            if ln is not None or gcn.relu or res_ok:
                xi = ops.layernorm_act(xi, None if ln is None else ln.weight, None if ln is None else ln.bias,
                                       relu=gcn.relu, residual=x if res_ok else None, out=xi)
            x = xi
        else:
            widths = [d] + [layer.att.heads * layer.att.out_channels for layer in model.att_layers]
            packed = torch.zeros((per, sum(widths)), dtype=torch.float32, device=dev)
            if rows > 0:
                xi = torch.empty((n, d), dtype=torch.float32, device=dev)
                conv(x, adj, out=xi, row0=r0, rows=rows)
                sh = xi[r0:r0 + rows]
                if ln is not None or gcn.relu or res_ok:
                    ops.layernorm_act(sh, None if ln is None else ln.weight, None if ln is None else ln.bias,
                                      relu=gcn.relu, residual=x[r0:r0 + rows] if res_ok else None, out=sh)
                xs = packed[:rows, :d]
                ops.layernorm_act(sh, model.gnn_norm.weight, model.gnn_norm.bias, relu=False, out=xs)
                c0 = d
                for layer, w in zip(model.att_layers, widths[1:]):
                    ops.linear(xs, layer.att.lin_r.weight[:, :d], out=packed[:rows, c0:c0 + w])
                    c0 += w
            full = torch.empty((per * world, sum(widths)), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(full, packed, group=group)          # the one collective of the eval
            full = full[:n]
            X = full[:, :d]
            kvs, c0 = [], d
            for w in widths[1:]:
                kvs.append(full[:, c0:c0 + w])
                c0 += w
            model._prime_kv(X, kvs)
            return X


def shard_queries(num_queries, rank, world):
    """Contiguous query-group shard [lo, hi) of this rank (a citation2 source's 1000 negatives stay together)."""
    per = (num_queries + world - 1) // world
    lo = min(num_queries, rank * per)
    return lo, min(num_queries, lo + per)


@torch.no_grad()
def score_links_sharded(model, score_func, links, X, test_set=False, group=None, group_size=1, batch_size=None):
    """Scores `links` [2, L] (L a multiple of group_size) with link groups sharded over the ranks of
    `group`; returns the full [L] probabilities on every rank (one all_gather of fp32 scores at the end)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = model._dev()
    L = links.shape[1]
    ngroups = L // group_size
    lo, hi = shard_queries(ngroups, rank, world)
    mine = links[:, lo * group_size:hi * group_size].to(dev)
    bs = batch_size or max(1, mine.shape[1])
    outs = [model.score_links(mine[:, s:s + bs], X, score_func, test_set=test_set) for s in range(0, mine.shape[1], bs)]
    local = torch.cat(outs) if outs else torch.empty(0, device=dev)
    if world == 1:
        return local
    per = (ngroups + world - 1) // world * group_size
    pad = torch.zeros(per, dtype=torch.float32, device=dev)
    pad[:local.numel()] = local
    full = torch.empty(per * world, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(full, pad, group=group)
    pieces = []
    for r in range(world):
        a, b = shard_queries(ngroups, r, world)
        pieces.append(full[r * per:r * per + (b - a) * group_size])
    return torch.cat(pieces)
