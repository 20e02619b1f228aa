"""Tiny driver for ncu: runs each hot kernel a few times on bench-sized inputs (no other work)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oov_b200
from oov_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda:0"
torch.manual_seed(0)
if which in ("score", "all"):
    Q, N, D, k = 1024, 1_000_000, 64, 20
    users = (torch.randn(Q, D, device=dev) * 0.3).to(torch.bfloat16)
    items = (torch.randn(N, D, device=dev) * 0.3).to(torch.bfloat16)
    hu = torch.randint(0, Q, (25_000,), device=dev)
    hi = torch.randint(1, N, (25_000,), device=dev)
    hist = ops.pairs_to_csr(hu, hi, Q)
    for _ in range(3):
        ops.fullsort_topk(users, items, k, hist=hist)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.fullsort_topk(users, items, k, hist=hist)
    e1.record(); torch.cuda.synchronize()
    print("score_topk ms", e0.elapsed_time(e1) / 5)
if which in ("linear", "all"):
    M, H = 1 << 18, 512
    A = torch.randn(M, H, device=dev).to(torch.bfloat16)
    W = torch.randn(H, H, device=dev).to(torch.bfloat16)
    b = torch.zeros(H, device=dev)
    for _ in range(3):
        ops.tc_linear(A, W, b, act="gelu", out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.tc_linear(A, W, b, act="gelu", out_dtype=torch.bfloat16)
    e1.record(); torch.cuda.synchronize()
    print("tc_linear ms", e0.elapsed_time(e1) / 5)
    for _ in range(3):
        ops.tc_linear(A, W, b, act="none", out_dtype=torch.bfloat16)
    e0.record()
    for _ in range(5):
        ops.tc_linear(A, W, b, act="none", out_dtype=torch.bfloat16)
    e1.record(); torch.cuda.synchronize()
    print("tc_linear(no act) ms", e0.elapsed_time(e1) / 5)
if which in ("hash", "all"):
    ids = torch.arange(500_000, 1_000_000, device=dev)
    keys = torch.randint(0, 256, (128, 16), device=dev, dtype=torch.uint8)
    for _ in range(3):
        ops.dhe_hash(ids, keys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.dhe_hash(ids, keys)
    e1.record(); torch.cuda.synchronize()
    print("dhe_hash ms", e0.elapsed_time(e1) / 5)
