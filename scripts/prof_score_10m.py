"""Score + top-k at 10 M items for a few pre-pass strides (OOV_SCORE_STRIDE is read once per process)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
D, k, Q, N = 64, 20, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
users = (torch.randn(Q, D, device=dev) * 0.3).to(torch.bfloat16)
items = (torch.randn(N, D, device=dev) * 0.3).to(torch.bfloat16)
hu = torch.randint(0, Q, (25 * Q,), device=dev); hi = torch.randint(1, N, (25 * Q,), device=dev)
hist = ops.pairs_to_csr(hu, hi, Q)
for _ in range(3): ops.fullsort_topk(users, items, k, hist=hist)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.fullsort_topk(users, items, k, hist=hist)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"stride {os.environ.get('OOV_SCORE_STRIDE', 'auto')} N={N}: {ms:.3f} ms  {2.0 * Q * N * D / ms / 1e9:.0f} TFLOP/s")
