"""Context-model rows of the scope table: token gather + OOV overwrite, the
first-order sum, the slsh / mean / zero embedders — achieved bytes/s against the algorithmic bytes of SURVEY 8d."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oov_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
B, fields = 65536, 26
g = np.random.default_rng(0)
vocab = np.maximum(10, np.exp(g.uniform(np.log(10), np.log(1e6), fields)).astype(np.int64))
vocab[0], vocab[1] = 1_000_000, 1_000_000                     # user_id, item_id
offsets = np.concatenate([[0], np.cumsum(vocab)[:-1]]).astype(np.int64)
V = int(vocab.sum())
tok = np.stack([np.minimum((g.zipf(1.1, B) - 1) % v, v - 1) for v in vocab], axis=1).astype(np.int64)
n_users = n_items = 800_000                                    # 20 % of ids 0/1 fall in the OOV range
tok[:, 0] = g.integers(0, 1_000_000, B); tok[:, 1] = g.integers(0, 1_000_000, B)
tokens, offs = torch.from_numpy(tok).to(dev), torch.from_numpy(offsets).to(dev)

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for D, name in ((16, "config 3 (DCNV2, D=16)"), (10, "config 4 (WideDeep/xDeepFM, D=10)")):
    for dt, s in ((torch.float32, 4), (torch.bfloat16, 2)):
        table = torch.randn(V, D, device=dev).to(dt)
        uc = torch.zeros(D, device=dev); ic = torch.zeros(D, device=dev)
        out = torch.empty((B, fields, D), dtype=dt, device=dev)
        ms = t(lambda: ops.token_gather(tokens, offs, table, n_users, n_items, user_const=uc, item_const=ic, out=out))
        alg = B * fields * (8 + 2 * s * D)
        print(f"token_gather {name} {str(dt)[6:]}: {ms * 1e3:.1f} us  algorithmic {alg / 1e6:.1f} MB -> {alg / ms / 1e6:.0f} GB/s ({100 * alg / ms / 1e6 / 6544.7:.0f} % of HBM)")
table1 = torch.randn(V, 1, device=dev)
ov = torch.zeros(1, device=dev)
ms = t(lambda: ops.first_order_sum(tokens, offs, table1, n_users, n_items, oov_user_val=ov, oov_item_val=ov))
alg = B * fields * (8 + 4) + B * 4
print(f"first_order_sum: {ms * 1e3:.1f} us  algorithmic {alg / 1e6:.1f} MB -> {alg / ms / 1e6:.0f} GB/s ({100 * alg / ms / 1e6 / 6544.7:.0f} % of HBM)")
# slsh / mean / zero on 1M OOV ids (D = 64)
n, F, D = 1_000_000, 32, 64
feat = torch.nn.functional.normalize(torch.randn(n, F, device=dev), dim=-1)
planes = torch.randn(10, F, device=dev)
W = torch.randn(1000, D, device=dev)
ids = torch.arange(n, device=dev)
for dt, s in ((torch.float32, 4), (torch.bfloat16, 2)):
    out = torch.empty((n, D), dtype=dt, device=dev)
    Wd = W.to(dt)
    ms = t(lambda: ops.slsh_embed(feat, planes, 1000, Wd, ids, out=out, out_dtype=dt), reps=5)
    alg = n * (8 + 4 * F + s * D + s * D)
    print(f"slsh_embed {str(dt)[6:]} n=1M: {ms * 1e3:.1f} us  algorithmic {alg / 1e6:.0f} MB -> {alg / ms / 1e6:.0f} GB/s ({100 * alg / ms / 1e6 / 6544.7:.0f} % of HBM)")
    vec = torch.randn(D, device=dev)
    ms = t(lambda: ops.const_embed(vec, ids, D, out=out, out_dtype=dt), reps=5)
    alg = n * (8 + s * D)
    print(f"const_embed (mean/zero) {str(dt)[6:]} n=1M: {ms * 1e3:.1f} us  algorithmic {alg / 1e6:.0f} MB -> {alg / ms / 1e6:.0f} GB/s ({100 * alg / ms / 1e6 / 6544.7:.0f} % of HBM)")
tbl = torch.randn(1_000_000, D, device=dev)
ms = t(lambda: ops.col_mean(tbl), reps=5)
print(f"col_mean 1M x 64 fp32: {ms * 1e3:.1f} us -> {tbl.numel() * 4 / ms / 1e6:.0f} GB/s ({100 * tbl.numel() * 4 / ms / 1e6 / 6544.7:.0f} % of HBM)")
ids2 = torch.randint(0, 2_000_000, (4_000_000,), device=dev)
ms = t(lambda: ops.map_ids(ids2, 1_000_000, 1000, "3round"), reps=5)
print(f"map_ids 4M (3round): {ms * 1e3:.1f} us -> {ids2.numel() * 16 / ms / 1e6:.0f} GB/s ({100 * ids2.numel() * 16 / ms / 1e6 / 6544.7:.0f} % of HBM)")
