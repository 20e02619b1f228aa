"""Fixed costs at the per-GPU sizes of an 8-way item shard (dhe1m: 62.5k in-vocab + 62.5k OOV rows per rank)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
D, k, Q = 64, 20, 1024
users = (torch.randn(Q, D, device=dev) * 0.3).to(torch.bfloat16)
hu = torch.randint(0, Q, (25 * Q,), device=dev); hi = torch.randint(1, 1_000_000, (25 * Q,), device=dev)
hist = ops.pairs_to_csr(hu, hi, Q)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for N in (62_500, 125_000, 250_000, 500_000, 1_250_000):
    items = (torch.randn(N, D, device=dev) * 0.3).to(torch.bfloat16)
    print(f"fullsort_topk Q={Q} N={N}: {t(lambda: ops.fullsort_topk(users, items, k, item_id_offset=500_000, hist=hist)):.3f} ms")
cs = torch.randn(16, Q, k, device=dev); ci = torch.randint(0, 1_000_000, (16, Q, k), device=dev)
print(f"topk_merge 16 lists: {t(lambda: ops.topk_merge(cs, ci)):.3f} ms")
