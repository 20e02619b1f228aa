"""Timing sweep of the fused score + mask + top-k kernel (CUDA events, inputs resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
D, k = 64, 20
cfgs = [(1024, 1_000_000), (4096, 1_000_000), (256, 1_000_000), (64, 1_000_000), (1, 1_000_000), (1024, 10_000_000)]
if len(sys.argv) > 1 and sys.argv[1] == "one":
    cfgs = cfgs[:1]
for Q, N in cfgs:
    users = (torch.randn(Q, D, device=dev) * 0.3).to(torch.bfloat16)
    items = (torch.randn(N, D, device=dev) * 0.3).to(torch.bfloat16)
    nh = 25 * Q
    hu = torch.randint(0, Q, (nh,), device=dev)
    hi = torch.randint(1, N, (nh,), device=dev)
    hist = ops.pairs_to_csr(hu, hi, Q)
    for name, h in (("hist", hist), ("nohist", None)):
        for _ in range(3):
            ops.fullsort_topk(users, items, k, hist=h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.fullsort_topk(users, items, k, hist=h)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        tf = 2.0 * Q * N * D / (ms * 1e-3) / 1e12
        gbs = (N * D * 2) / (ms * 1e-3) / 1e9
        print(f"score_topk Q={Q} N={N} {name}: {ms:.3f} ms  {tf:.1f} TFLOP/s  item-table {gbs:.0f} GB/s  {Q / (ms * 1e-3):.0f} queries/s")
