"""Timeline of CTA 0 of tc_lsh_embed_kernel (clock64 stamps written through the OOV_LSH_TRACE_PTR profiling hook)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
n = 148 * 128 * 12
F, B, D = 32, 1000, 64
feat = torch.nn.functional.normalize(torch.randn(n, F, device=dev), dim=-1)
planes = torch.randn(B, F, device=dev)
W = torch.randn(B, D, device=dev) * 0.1
ids = torch.arange(n, device=dev)
out = torch.empty((n, D), dtype=torch.bfloat16, device=dev)
ops.lsh_embed(feat, planes, W, ids, out=out, n_old=0, path=ops.PATH_TCGEN05)
trace = torch.zeros(4 * 4096, dtype=torch.int64, device=dev)
os.environ["OOV_LSH_TRACE_PTR"] = hex(trace.data_ptr())
ops.lsh_embed(feat, planes, W, ids, out=out, n_old=0, path=ops.PATH_TCGEN05)
torch.cuda.synchronize()
del os.environ["OOV_LSH_TRACE_PTR"]
tr = trace.cpu().view(4, 4096)
names = {0: "m1.tile", 1: "m1.wait", 2: "m1.go", 3: "m1.issued", 4: "m2.wait", 5: "m2.go", 6: "m2.issued", 7: "w.wait_acc1", 8: "w.acc1_ready",
         9: "w.loaded", 10: "w.signs_done", 11: "w.h_empty", 12: "w.stored", 13: "w.staged_next", 14: "w.queue_done", 15: "w.acc2_ready", 16: "w.tile_done", 17: "w.features_loaded", 18: "w.stage_barrier", 19: "w.a_empty"}
ev = []
for role in range(4):
    for x in tr[role].tolist():
        if x == 0:
            continue
        x &= (1 << 64) - 1
        ev.append((x & ((1 << 48) - 1), role, x >> 56, (x >> 48) & 255))
ev.sort()
t0 = ev[0][0]
lim = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for t, role, e, tag in ev[:lim]:
    print(f"{t - t0:9d}  {'  ' * role * 6}{['MMA1', 'MMA2', 'W0', 'W15'][role]} {names.get(e, e)} #{tag}")
