"""Hit chunks per CTA of the column-split main pass (OOV_SCORE_DEBUG=32) on the bench workloads' own data."""
import os, sys
os.environ["OOV_SCORE_DEBUG"] = "32"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import numpy as np
import torch
from oov_b200 import ops
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
for name in sys.argv[1:] or ["dhe1m"]:
    wl = dict(b.WORKLOADS[name])
    cfg, emb, model = b.build_gpu(wl, "cuda:0", 0)
    N, k, Q = wl["n_items"], wl["k"], wl["Q"]
    table = model.build_item_table(N)
    users, hu, hi = b.query_batch(wl, 100)
    u = torch.from_numpy(users).cuda()
    csr = ops.pairs_to_csr(torch.from_numpy(hu).cuda(), torch.from_numpy(hi).cuda(), Q)
    user_e = model._assemble("user", u, out_dtype=model.table_dtype)
    ops.fullsort_topk(user_e, table, k, hist=csr)
    torch.cuda.synchronize()
    ws = max(ops._ws_cache.values(), key=lambda t: t.numel())
    cnt = ws.view(torch.uint8)[: 148 * 4].cpu().numpy().view(np.uint32)
    print(name, "hit chunks per CTA: total", int(cnt.sum()), "max", int(cnt.max()), "per CTA:", cnt.reshape(2, 74).tolist())
    ue = user_e.float()
    print("  user_e abs mean old/new:", ue[u < wl["n_old_users"]].abs().mean().item(), ue[u >= wl["n_old_users"]].abs().mean().item(),
          " table abs mean old/new:", table[: wl["n_old_items"]].float().abs().mean().item(), table[wl["n_old_items"]:].float().abs().mean().item())
    del model, emb, table
    torch.cuda.empty_cache()
