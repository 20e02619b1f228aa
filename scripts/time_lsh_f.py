"""tc_lsh_embed timing for F = 32 / 48 / 64 against the exact-sign CUDA-core path (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
n, B, D = 2_000_000, 1000, 64
g = torch.Generator(device="cpu").manual_seed(0)
for F in (32, 48, 64):
    feat = torch.nn.functional.normalize(torch.randn(n, F, generator=g), dim=-1).to(dev)
    planes = torch.randn(B, F, generator=g).to(dev)
    W = (torch.randn(B, D, generator=g) * 0.1).to(dev)
    ids = torch.arange(n, device=dev)
    out = torch.empty((n, D), dtype=torch.bfloat16, device=dev)
    for path, name in ((ops.PATH_TCGEN05, "tcgen05"), (ops.PATH_SIMT_FP32, "simt fp32")):
        nn = n if path == ops.PATH_TCGEN05 else n // 10
        f = lambda: ops.lsh_embed(feat, planes, W, ids[:nn], out=out[:nn], n_old=0, path=path)
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"F={F} {name}: {nn} ids {ms:.3f} ms = {nn / ms / 1e6:.3f} G ids/s")
