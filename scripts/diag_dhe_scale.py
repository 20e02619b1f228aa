"""Diagnostic: DHE product rows vs oracle for growing n / id offsets (which size breaks?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import numpy as np, torch
from oracle import oracle as o
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
from oov_b200 import ops
dev = "cuda:0"
H, hid, D = 128, 512, 64
keys_b = b.dhe_keys(H)
keys = o.keys_to_array(keys_b)
ws, bs = b.dhe_weights(H, hid, D, 5)
t = lambda a: torch.from_numpy(a).to(dev)
net = ops.DheNet([t(w) for w in ws], [t(x) for x in bs])
keys_dev = ops.keys_tensor(keys_b, dev)
w16 = [o.round_bf16(w) for w in ws]
for lo, n in ((300, 40_000), (500_000, 40_000), (500_000, 200_000), (500_000, 262_144), (500_000, 300_000), (500_000, 500_000), (0, 600_000)):
    ids = torch.arange(lo, lo + n, device=dev)
    for path, nm in ((ops.PATH_TCGEN05, "tc"), (ops.PATH_SIMT_FP32, "simt")):
        if nm == "simt" and n > 300_000:
            continue
        out = ops.dhe_embed(ids, keys_dev, net, out_dtype=torch.bfloat16, path=path).float().cpu().numpy()
        sel = np.unique(np.concatenate([np.arange(0, 64), np.arange(n // 2, n // 2 + 64), np.arange(n - 64, n), np.random.default_rng(0).integers(0, n, 256)]))
        want = o.dhe_mlp(o.dhe_hashes(np.arange(lo, lo + n, dtype=np.int64)[sel], keys), w16, bs, bf16_points=True)
        err = np.abs(out[sel] - want) / np.maximum(np.abs(want), 1e-6)
        bad = np.nonzero(err.max(axis=1) > 1e-2)[0]
        print(f"lo={lo} n={n} {nm}: max rel {err.max():.3e}  bad rows {len(bad)}/{len(sel)} first bad sel idx {sel[bad[:5]] if len(bad) else ''}", flush=True)
