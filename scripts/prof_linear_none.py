import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
M = 1 << 18
for (N, K, act) in ((512, 64, "none"), (512, 512, "none"), (512, 512, "gelu")):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = torch.randn(N, K, device=dev).to(torch.bfloat16)
    b = torch.zeros(N, device=dev)
    for _ in range(3):
        ops.tc_linear(A, W, b, act=act, out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.tc_linear(A, W, b, act=act, out_dtype=torch.bfloat16)
    e1.record(); torch.cuda.synchronize()
    print(N, K, act, "ms", e0.elapsed_time(e1) / 5)
