"""One rank's share of the sharded dhe1m step (rank 0 of 8, no collective), eager vs captured in a CUDA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oov_b200 import ops, sharded

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wlname = sys.argv[2] if len(sys.argv) > 2 else "dhe1m"
wl = dict(bench.WORKLOADS[wlname])
dev = "cuda:0"
torch.cuda.set_device(0)
cfg, emb, model = bench.build_gpu(wl, dev, 0)
Q, k, N = wl["Q"], wl["k"], wl["n_items"]
users, hu, hi = bench.query_batch(wl, 100)
u = torch.from_numpy(users).to(dev)
dhu, dhi = torch.from_numpy(hu).to(dev), torch.from_numpy(hi).to(dev)
csr = ops.pairs_to_csr(dhu, dhi, Q)
sr = sharded.ShardedRetrieval(model, N, rank=0, world_size=world)
sr.world = 1                     # no process group here: skip the all-gather, keep the 1/world segments
print("segments", sr.segments)

def step():
    user_e = model._assemble("user", u, out_dtype=model.table_dtype)
    sr.build_shard()
    return sr.topk(user_e, k, hist_pairs=(dhu, dhi))

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (t1 - t0) * 1e3 / reps

g_ms, c_ms = timed(step)
print(f"eager : GPU {g_ms:.3f} ms/step, CPU enqueue {c_ms:.3f} ms/step")
for name, fn in (("user_embed", lambda: model._assemble("user", u, out_dtype=model.table_dtype)),
                 ("build_shard", sr.build_shard)):
    g, c = timed(fn)
    print(f"  {name}: GPU {g:.3f} ms, CPU {c:.3f} ms")
ue = model._assemble("user", u, out_dtype=model.table_dtype)
g, c = timed(lambda: sr.topk(ue, k, hist_pairs=(dhu, dhi)))
print(f"  topk (CSR build + one fused launch over both segments + merge): GPU {g:.3f} ms, CPU {c:.3f} ms")
g, c = timed(lambda: sr.topk(ue, k, hist=csr))
print(f"  topk (per-segment path, prebuilt CSR): GPU {g:.3f} ms, CPU {c:.3f} ms")

import oov_b200
gq = oov_b200.GraphedTopK(model, Q, k, N, Q * wl["max_hist"], sharded=sr)
out_s, out_i = gq(u, dhu, dhi)
torch.cuda.synchronize()
ref_s, ref_i = step()
torch.cuda.synchronize()
print("GraphedTopK == eager:", torch.equal(out_i, ref_i), torch.equal(out_s, ref_s))
g_ms, c_ms = timed(lambda: gq(u, dhu, dhi))
print(f"GraphedTopK (query side forked onto a second stream): GPU {g_ms:.3f} ms/step, CPU enqueue {c_ms:.3f} ms/step")
