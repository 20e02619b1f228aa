"""How the sampled hits of the scoring pre-pass coincide across users, on the bench workloads' own tables (torch maths)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import numpy as np
import torch
from oov_b200 import ops
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
for name, stride in (("lsh1m", 4), ("dhe1m", 4)):
    wl = dict(b.WORKLOADS[name])
    cfg, emb, model = b.build_gpu(wl, "cuda:0", 0)
    N, k, Q = wl["n_items"], wl["k"], wl["Q"]
    table = model.build_item_table(N)
    users, hu, hi = b.query_batch(wl, 100)
    ue = model._assemble("user", torch.from_numpy(users).cuda(), out_dtype=model.table_dtype).float()
    nt = N // 128
    tiles = torch.arange(0, nt, stride, device="cuda")
    tmax = torch.empty((Q, tiles.numel()), device="cuda")
    for c0 in range(0, tiles.numel(), 256):
        tt = tiles[c0:c0 + 256]
        rows = (tt[:, None] * 128 + torch.arange(128, device="cuda")[None, :]).reshape(-1)
        s = ue @ table[rows].float().T
        tmax[:, c0:c0 + 256] = s.reshape(Q, -1, 128).amax(dim=2)
    tmax = torch.nan_to_num(tmax, nan=float("inf"))
    R = k + 3
    T = torch.topk(tmax, R, dim=1).values[:, -1]
    hits = (tmax >= T[:, None]).sum(0).float()
    srt = torch.sort(hits, descending=True).values
    print(name, "sampled tiles", tiles.numel(), "mean", hits.mean().item(), "max", srt[0].item(), "top-8", srt[:8].tolist(),
          "p99", srt[int(0.01 * srt.numel())].item(), "tiles > 6 x mean:", int((hits > 6 * hits.mean()).sum()),
          "second moment / mean", (hits * hits).sum().item() / hits.sum().item())
    del model, emb, table
    torch.cuda.empty_cache()
