"""Kernel timings of the DCN-V2 tower pieces (CUDA events, graph replays)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oov_b200
from oov_b200 import ops

dev = "cuda:0"
M = 65536


def t(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        k = fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for N, K in ((416, 416), (512, 416), (256, 416), (448, 448), (512, 512), (768, 416), (768, 768), (384, 384)):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    us = t(lambda: ops.tc_linear(A, W, b, act="none", out_dtype=torch.bfloat16))
    us32 = t(lambda: ops.tc_linear(A, W, b, act="none", out_dtype=torch.float32))
    line = f"M={M} N={N} K={K}: tc_linear bf16 {us:7.1f} us ({2.0*M*N*K/us/1e6:6.0f} TF/s)  fp32-out {us32:7.1f} us"
    if N == K:
        x0 = torch.randn(M, N, device=dev).to(torch.bfloat16)
        usu = t(lambda: ops.cross_update(x0, A, A))
        line += f"  cross_update {usu:6.1f} us"
    print(line)
