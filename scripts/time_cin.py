"""xDeepFM CIN at the bench shape (65536 rows x 26 fields x D 10, CIN 100-100-100): fused kernel vs the three-launch path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oov_b200
from oov_b200 import ops
from oov_b200.inductive.zero_embedder import ZeroEmbedder
dev = "cuda:0"
Bn, fields, D = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 26, 10
cfg = {"embedding_size": D, "mlp_hidden_size": [128, 128, 128], "dropout_prob": 0.2, "device": dev, "direct": False, "cin_layer_size": [100, 100, 100]}
z = lambda d: ZeroEmbedder(np.zeros((10, 1), np.float32), np.zeros((10, 1), np.float32), 40, 40, d, dev)
torch.manual_seed(0)
m = oov_b200.xDeepFM(cfg, [40, 40] + [50] * (fields - 2), inductive_embedder=z(D), first_order_embedder=z(1)).to(dev).eval()
m.pack_tower()
x = (torch.randn(Bn, fields, D, device=dev) * 0.3).to(torch.bfloat16)

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

hs = [26, 50, 50]
flops = 2.0 * Bn * D * sum(h * 26 * 100 for h in hs)
for fused in (True, False):
    m.fused_cin = fused
    ms = timed(lambda: m.compressed_interaction_network(x))
    print(f"fused={fused}: CIN {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s (algorithmic 2 B D sum H M O)")
