import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
D = 64
Q, N = 1024, 1_000_000
users = (torch.randn(Q, D, device=dev) * 0.3).to(torch.bfloat16)
items = (torch.randn(N, D, device=dev) * 0.3).to(torch.bfloat16)
for k in (1, 5, 10, 20, 24):
    for _ in range(3):
        ops.fullsort_topk(users, items, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.fullsort_topk(users, items, k)
    e1.record(); torch.cuda.synchronize()
    print(f"k={k}: {e0.elapsed_time(e1) / 5:.3f} ms")
