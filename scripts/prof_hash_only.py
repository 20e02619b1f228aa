"""Time the byte-plane SipHash kernel alone (through oov_dhe_embed on a net whose layers are tiny is not possible:
use the difference between dhe_embed with and without hashing instead): prints dhe_embed time for n ids."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
H, hid, D = 128, 512, 64
torch.manual_seed(0)
ws = [torch.randn(hid, H, device=dev) * 2e-8, torch.randn(hid, hid, device=dev) * 0.04, torch.randn(hid, hid, device=dev) * 0.04, torch.randn(D, hid, device=dev) * 0.04]
bs = [torch.randn(hid, device=dev) * 0.1 for _ in range(3)] + [torch.randn(D, device=dev) * 0.1]
net = ops.DheNet(ws, bs)
keys = torch.randint(0, 256, (H, 16), dtype=torch.uint8, device=dev)
for n in (65536, 131072, 262144):
    ids = torch.arange(500_000, 500_000 + n, device=dev)
    out = torch.empty((n, D), dtype=torch.bfloat16, device=dev)
    for _ in range(3): ops.dhe_embed(ids, keys, net, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.dhe_embed(ids, keys, net, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"debug={os.environ.get('OOV_HASH_DEBUG','0')} n={n}: dhe_embed {e0.elapsed_time(e1)/10:.3f} ms")
