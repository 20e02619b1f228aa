"""Timing of the LSH OOV embed (tensor-core path vs CUDA-core path), inputs resident."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
F, B, D = 32, 1000, 64
feat = torch.nn.functional.normalize(torch.randn(n, F, device=dev), dim=-1)
planes = torch.randn(B, F, device=dev)
W = torch.randn(B, D, device=dev) * 0.1
ids = torch.arange(n, device=dev)
paths = (("tcgen05", ops.PATH_TCGEN05), ("simt", ops.PATH_SIMT_FP32)) if n <= 1_000_000 else (("tcgen05", ops.PATH_TCGEN05),)
for name, path in paths:
    out = torch.empty((n, D), dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        ops.lsh_embed(feat, planes, W, ids, out=out, n_old=0, path=path)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.lsh_embed(feat, planes, W, ids, out=out, n_old=0, path=path)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    dense = 2.0 * n * B * (F + D)
    issued = 2.0 * n * 1024 * (64 + 64)
    print(f"lsh_embed {name} n={n}: {ms:.3f} ms  {n / ms / 1e3:.2f} M ids/s  dense-equivalent {dense / ms / 1e9:.1f} TFLOP/s  issued-MMA {issued / ms / 1e9:.1f} TFLOP/s")
