"""DHE item embedding (SipHash byte planes + 4 tcgen05 layers) timing; OOV_HASH_VARIANT selects the hash kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
H, hid, D = 128, 512, 64
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
ws = [torch.randn(hid, H, device=dev) * 2e-8, torch.randn(hid, hid, device=dev) * 0.04, torch.randn(hid, hid, device=dev) * 0.04,
      torch.randn(D, hid, device=dev) * 0.04]
bs = [torch.randn(hid, device=dev) * 0.1, torch.randn(hid, device=dev) * 0.1, torch.randn(hid, device=dev) * 0.1, torch.randn(D, device=dev) * 0.1]
net = ops.DheNet(ws, bs)
keys = torch.randint(0, 256, (H, 16), dtype=torch.uint8, device=dev)
ids = torch.arange(500_000, 500_000 + n, device=dev)
out = torch.empty((n, D), dtype=torch.bfloat16, device=dev)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: ops.dhe_embed(ids, keys, net, out=out))
chk = out.float().sum().item()
print(f"variant {os.environ.get('OOV_HASH_VARIANT', 'default')}: dhe_embed n={n}: {ms:.3f} ms  ({n / ms / 1e3:.1f} M ids/s)  checksum {chk:.6f}")
