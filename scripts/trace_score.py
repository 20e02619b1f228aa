"""Timeline of tc_score_main2_kernel built with -DOOV_SCORE_TRACE (events of CTA (0,0) in the workspace's first bytes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oov_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
D, k, Q, N = 64, 20, 1024, int(os.environ.get('N', 10_000_000))
users = (torch.randn(Q, D, device=dev) * 0.3).to(torch.bfloat16)
items = (torch.randn(N, D, device=dev) * 0.3).to(torch.bfloat16)
hu = torch.randint(0, Q, (25 * Q,), device=dev); hi = torch.randint(1, N, (25 * Q,), device=dev)
hist = ops.pairs_to_csr(hu, hi, Q)
ops.fullsort_topk(users, items, k, hist=hist)
torch.cuda.synchronize()
ws = [b for key, b in ops._ws_cache.items()]
ws = max(ws, key=lambda b: b.numel())
ws.view(torch.uint8)[: 25 * 2048 * 8].zero_()
ops.fullsort_topk(users, items, k, hist=hist)
torch.cuda.synchronize()
raw = ws.view(torch.uint8)[: 25 * 2048 * 8].cpu().numpy().view(np.uint64).reshape(25, 2048)
names = {0: "acc_empty ok", 1: "B full ok", 2: "mma issued", 3: "acc_full ok", 4: "ldtm done", 5: "stage empty ok"}
ev = []
for w in range(25):
    for x in raw[w]:
        x = int(x)
        if x == 0:
            continue
        ev.append((x >> 24, w, (x >> 20) & 15, (x >> 16) & 15, x & 0xffff))
ev.sort()
tmin = ev[0][0] if ev else 0
import collections
done = collections.defaultdict(dict)
for c, w, e, ut, t in ev:
    if e == 4:
        done[(t, ut)][w] = c - tmin
prev_first = None
for key in sorted(done):
    d = done[key]
    if len(d) < 16 or not (30 <= key[0] < 50):
        continue
    first = min(d.values())
    late = {w: v - first for w, v in d.items() if v - first > 350}
    print(key, "first", first, "" if prev_first is None else f"(+{first - prev_first})", "spread", max(d.values()) - first, "late warps", late)
    prev_first = first
