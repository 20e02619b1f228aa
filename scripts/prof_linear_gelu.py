"""One DHE hidden layer (262144 x 512 -> 512, GELU, bf16) through oov_tc_linear: timing, or the ncu target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
M, N, K = 1 << 18, 512, 512
A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
W = (torch.randn(N, K, device=dev) * 0.04).to(torch.bfloat16)
b = torch.randn(N, device=dev) * 0.1
for _ in range(3):
    ops.tc_linear(A, W, b, act="gelu", out_dtype=torch.bfloat16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.tc_linear(A, W, b, act="gelu", out_dtype=torch.bfloat16)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"tc_linear GELU M={M} N={N} K={K}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
