"""world_size-2 gloo test of the link-sharding host logic (no GPU): every link is scored exactly once, by
the rank that owns its query group, and the gathered result is in the original order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _StubModel:
    """Stands in for LinkTransformer: the 'score' of link (a,b) is a deterministic function of (a,b),
    tagged with the rank that computed it."""

    def __init__(self, rank):
        self.rank = rank
        self.seen = 0

    def _dev(self):
        return torch.device("cpu")

    def score_links(self, batch, X, score_func, test_set=False):
        self.seen += batch.shape[1]
        return (batch[0] * 1000 + batch[1]).float() + 0.25 * self.rank


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lpformer_b200.evaluate import score_links_sharded, shard_queries
    group_size, nq = 5, 7                                  # 7 query groups of 5 links: uneven over 2 ranks
    links = torch.stack([torch.arange(nq).repeat_interleave(group_size), torch.arange(nq * group_size) % 11])
    model = _StubModel(rank)
    out = score_links_sharded(model, None, links, None, group_size=group_size, batch_size=4)
    lo, hi = shard_queries(nq, rank, world)
    assert model.seen == (hi - lo) * group_size
    expect = (links[0] * 1000 + links[1]).float()
    owner = torch.zeros(nq * group_size)
    for r in range(world):
        a, b = shard_queries(nq, r, world)
        owner[a * group_size:b * group_size] = r
    assert torch.equal(out, expect + 0.25 * owner)
    ret[rank] = True
    dist.barrier()
    dist.destroy_process_group()


def test_link_sharding_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_shard_queries_partition():
    from lpformer_b200.evaluate import shard_queries
    for n in (0, 1, 7, 64, 1001):
        for w in (1, 2, 3, 8):
            spans = [shard_queries(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_shard_bounds_balance_by_cost():
    """shard_bounds with a cost per query group (balance by the weight of the links' rows, not by their
    number): contiguous, complete, monotone, and every shard within one group's cost of the ideal share."""
    from lpformer_b200.evaluate import shard_bounds, shard_queries
    g = torch.Generator().manual_seed(3)
    for n in (1, 7, 64, 1001):
        for w in (1, 2, 3, 8):
            assert shard_bounds(n, w) == [shard_queries(n, r, w)[0] for r in range(w)] + [n]
            cost = torch.rand(n, generator=g, dtype=torch.float64) ** 4 * 1000 + 1      # a few heavy groups
            cost[n // 2] = 5000.0
            b = shard_bounds(n, w, cost)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            share = float(cost.sum()) / w
            for r in range(w):
                got = float(cost[b[r]:b[r + 1]].sum())
                assert got <= share + float(cost.max()) + 1e-6, (n, w, r, got, share)
    assert shard_bounds(0, 4, torch.zeros(0)) == [0, 0, 0, 0, 0]
    assert shard_bounds(5, 2, torch.zeros(5)) == [0, 3, 5]                              # zero cost: by count
