"""LSH OOV embed of 5 M ids with the fp32 -> bf16 cast of 5 M in-vocab rows folded into the launch (oov_lsh_embed_cast),
against the two as separate launches; inputs resident.  Driver of the ncu capture profiles/r02_lsh_cast_ncu.txt."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
F, B, D = 32, 1000, 64
feat = torch.nn.functional.normalize(torch.randn(n, F, device=dev), dim=-1)
planes = torch.randn(B, F, device=dev)
W = torch.randn(B, D, device=dev) * 0.1
ids = torch.arange(n, device=dev)
table = torch.randn(n, D, device=dev) * 0.1                      # the in-vocab rows (fp32 item_embedding.weight)
out = torch.empty((2 * n, D), dtype=torch.bfloat16, device=dev)


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


fused = timed(lambda: ops.lsh_embed(feat, planes, W, ids, out=out[n:], n_old=0, path=ops.PATH_TCGEN05, side_cast=(table, out[:n])))
ref = out.clone()
alone = timed(lambda: ops.lsh_embed(feat, planes, W, ids, out=out[n:], n_old=0, path=ops.PATH_TCGEN05))
cast = timed(lambda: ops.gather_rows(table, ids, out=out[:n]))
print(f"n={n}: LSH + cast in one launch {fused:.3f} ms | LSH alone {alone:.3f} ms + gather-cast launch {cast:.3f} ms = {alone + cast:.3f} ms | same table: {torch.equal(ref.view(torch.int16), out.view(torch.int16))}")
