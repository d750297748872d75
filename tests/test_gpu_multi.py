"""-m gpu, needs >= 2 devices (skipped otherwise): the multi-GPU eval path under NCCL — row-sharded last GCN layer +
all-gather of the packed [X | KV] tables (evaluate.propagate_replicated) and link-sharded scoring
(evaluate.evaluate_mrr's sharding) — against the single-GPU results."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import lpformer_b200 as L
    from lpformer_b200 import synthetic as S
    from lpformer_b200.evaluate import LinkScoreStream, propagate_replicated, shard_queries
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = S.make_graph("citation2", seed=21, scale=0.02, heldout=128)
    torch.manual_seed(3)
    model = L.LinkTransformer(S.train_args_of(g.cfg), g.data_dict(dev), device=dev).to(dev).eval()
    score = L.mlp_score(model.out_dim, model.out_dim, 1, 2).to(dev).eval()
    X_rep = propagate_replicated(model)              # sharded last layer + one all-gather
    X_one = model.propagate()                        # the whole GCN on this GPU
    nq, negs = 16, 100
    links = torch.from_numpy(S.citation2_queries(g, nq, negs, seed=5)).to(dev)
    q0, q1 = shard_queries(nq, rank, world)
    mine = links[:, q0 * (1 + negs): q1 * (1 + negs)]
    part = LinkScoreStream(model, score, X_rep, mine.shape[1], depth=2).score(mine)
    full = model.score_links(links, X_one, score)
    parts = [torch.empty(shard_queries(nq, r, world)[1] * (1 + negs) - shard_queries(nq, r, world)[0] * (1 + negs),
                         device=dev) for r in range(world)]
    dist.all_gather(parts, part.contiguous())
    if rank == 0:
        out["x_diff"] = float((X_rep - X_one).abs().max())
        out["score_diff"] = float((torch.cat(parts) - full).abs().max())
        out["checksum"] = float(torch.cat(parts).double().sum())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_eval_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert out["x_diff"] <= 1e-5, out["x_diff"]          # same kernels, row-sharded: identical up to fp32 re-association
        assert out["score_diff"] <= 1e-6, out["score_diff"]
        assert np.isfinite(out["checksum"])
