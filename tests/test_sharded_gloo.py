"""CPU, world_size = 2 over gloo: the exchange logic of the row-sharded retrieval (shard ownership,
one packed all-gather, deterministic merge).  The CUDA kernels cannot run here, so the two device ops
are replaced by oracle stand-ins (tests may use the oracle; the product default stays CUDA-only) —
what is under test is everything around them."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as o

N_OLD, N_ALL, D, Q, K = 90, 231, 16, 13, 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    g = np.random.default_rng(5)
    table = g.standard_normal((N_ALL, D)).astype(np.float32)
    table[100] = table[200]                               # a score tie that straddles the two ranks
    users = g.standard_normal((Q, D)).astype(np.float32)
    hu = g.integers(0, Q, 40)
    hi = g.integers(1, N_ALL, 40)
    return table, users, hu, hi


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oov_b200 import sharded
        table, users, hu, hi = _inputs()

        class FakeModel:
            n_items = N_OLD

        def build(lo, hi_):                               # stands in for model.build_item_table(row_range=...)
            return torch.from_numpy(table[lo:hi_].copy())

        def local_topk(ue, tab, k, off, seg, hist):       # stands in for ops.fullsort_topk on this shard
            s = o.full_sort_scores(ue.numpy(), tab.numpy())
            gids = np.arange(tab.shape[0]) + off
            s[:, gids == 0] = -np.inf
            for u, i in zip(hu, hi):
                if off <= i < off + tab.shape[0]:
                    s[u, i - off] = -np.inf
            v, idx = o.topk(s, k)
            pad = k - v.shape[1]
            v = np.pad(v, ((0, 0), (0, pad)), constant_values=-np.inf)
            gi = np.pad(idx + off, ((0, 0), (0, pad)), constant_values=-1)
            return torch.from_numpy(v), torch.from_numpy(gi)

        def merge(cs, ci):
            cs, ci = cs.numpy().copy(), ci.numpy().copy()
            cs[ci < 0] = -np.inf
            ci[ci < 0] = np.iinfo(np.int64).max
            v, i = o.merge_topk(cs, ci, cs.shape[2])
            return torch.from_numpy(v), torch.from_numpy(i)

        sr = sharded.ShardedRetrieval(FakeModel(), N_ALL, local_topk_fn=local_topk, merge_fn=merge, build_table_fn=build)
        assert sr.rank == rank and sr.world == world and len(sr.segments) == 2
        s, i = sr.topk(torch.from_numpy(users), K)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), s=s.numpy(), i=i.numpy(), seg=np.array(sr.segments))
    finally:
        dist.destroy_process_group()


def test_sharded_topk_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert (r0["i"] == r1["i"]).all() and (r0["s"] == r1["s"]).all()          # every rank holds the same answer
    # ownership: in-vocab and OOV ranges are each split across the ranks
    assert r0["seg"].tolist() == [[0, 45], [90, 161]] and r1["seg"].tolist() == [[45, 90], [161, 231]]
    table, users, hu, hi = _inputs()
    full = o.mask_scores(o.full_sort_scores(users, table), hu, hi)
    wv, wi = o.topk(full, K)
    assert (r0["i"] == wi).all()
    assert np.allclose(r0["s"], wv, rtol=0, atol=0)
