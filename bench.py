#!/usr/bin/env python
"""bench.py — full-sort top-k queries/s with LSH / DHE OOV embedding on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU arm (oracle port of the reference path)

A *step* is one pass of the hot path over one batch of Q query users, un-amortised like the reference
(bpr.py:151-156 re-embeds every item per batch): embed the Q users (in-vocab gather + OOV embed), embed ALL N
items (in-vocab gather + OOV embed) into the bf16 item table, score Q x N, mask pad + history, top-k.
Default workload `lsh10m` = BASELINE.json configs[4], the configuration the metric's target is quoted on (BPR + lsh,
10M items of which 5M OOV, F = 32, B = 1000, D = 64, Q = 1024, k = 20; it fits one B200).  The same JSON line carries
a second block `workloads.dhe1m` = configs[1] (DirectAU + dhe, 1M items / 100k users, bf16) with its own value / e2e /
roofline, measured in the same process (`--single` skips it), and at N = 1 the ranking blocks `workloads.dcnv2_criteo` =
configs[2] (DCNV2 tower with slsh OOV buckets, 65536 rows x 26 token fields per step, rows/s) and
`workloads.xdeepfm_criteo` = configs[3] (xDeepFM: first-order + CIN + MLP with the `mean` embedder, same rows).
With N > 1 the item rows are sharded over the ranks (strong scaling: total work fixed), each rank embeds and
scores its shard, one NCCL all-gather moves the [S, Q, k] candidates and every rank merges.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "dhe1m": dict(model="DirectAU", embedder="dhe", n_items=1_000_000, n_old_items=500_000, n_users=100_000,
                  n_old_users=50_000, D=64, H=128, hidden=512, Q=1024, k=20, max_hist=50),
    # BASELINE.json configs[4] (scale sweep)
    "lsh10m": dict(model="BPR", embedder="lsh", n_items=10_000_000, n_old_items=5_000_000, n_users=100_000,
                   n_old_users=50_000, D=64, F=32, B=1000, Q=1024, k=20, max_hist=50),
    # small variants for quick checks
    "dhe100k": dict(model="DirectAU", embedder="dhe", n_items=100_000, n_old_items=50_000, n_users=10_000,
                    n_old_users=5_000, D=64, H=128, hidden=512, Q=256, k=20, max_hist=50),
    "lsh1m": dict(model="BPR", embedder="lsh", n_items=1_000_000, n_old_items=500_000, n_users=100_000,
                  n_old_users=50_000, D=64, F=32, B=1000, Q=1024, k=20, max_hist=50),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lsh10m", choices=list(WORKLOADS))
    ap.add_argument("--second", default="dhe1m", choices=list(WORKLOADS), help="second workload reported under `workloads`")
    ap.add_argument("--single", action="store_true", help="measure only --workload")
    ap.add_argument("--no-dcnv2", dest="dcnv2", action="store_false", help="skip the workloads.dcnv2_criteo / xdeepfm_criteo blocks (N = 1 only)")
    ap.add_argument("--Q", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--cpu-slices", type=int, default=None, help="CPU arm: item-axis slices per query batch (one slice = one step)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------- synthetic inputs
def dhe_keys(H):
    g = np.random.Generator(np.random.PCG64(1234))
    return [bytes(r.tolist()) for r in g.integers(0, 256, size=(H, 16), dtype=np.uint8)]


def dhe_weights(H, hidden, D, seed):
    """nn.Linear-shaped nets with 'trained-looking' magnitudes (layer 1 scaled so activations are O(1))."""
    g = np.random.Generator(np.random.PCG64(seed))
    dims = [H, hidden, hidden, hidden, D]
    ws, bs = [], []
    for l in range(4):
        bound = 1.0 / np.sqrt(dims[l])
        w = g.uniform(-bound, bound, size=(dims[l + 1], dims[l])).astype(np.float32)
        if l == 0:
            w *= np.float32(2e-7)
        ws.append(w)
        bs.append(g.uniform(-bound, bound, size=(dims[l + 1],)).astype(np.float32))
    return ws, bs


def query_batch(wl, seed):
    """Q user ids (half in-vocab, half OOV) + history pairs (0..max_hist items per user)."""
    g = np.random.Generator(np.random.PCG64(seed))
    Q = wl["Q"]
    users = np.where(g.random(Q) < 0.5, g.integers(1, wl["n_old_users"], Q), g.integers(wl["n_old_users"], wl["n_users"], Q))
    counts = g.integers(0, wl["max_hist"] + 1, Q)
    hu = np.repeat(np.arange(Q), counts)
    hi = g.integers(1, wl["n_items"], hu.shape[0])
    return users.astype(np.int64), hu.astype(np.int64), hi.astype(np.int64)


# --------------------------------------------------------------------------------------- CPU arm (oracle port)
class CpuReference:
    """The reference's CPU path, restated by the oracle (the reference itself is Python on torch with absent
    dependencies and does not travel to the GPU box), at the FULL workload size: no sub-sampling, no scaled times.

    One query batch = embed the Q users, embed ALL N items, score, mask pad + history, top-k.  The reference
    materialises [N, B] multi-hot and [Q, N] score matrices (40 GB each at 10M items), so the item axis is walked in
    `slices` equal slices with a running top-k merge; ONE SLICE IS ONE STEP, `slices` steps complete one query batch:
    queries/s = Q / (slices x seconds per step), every number is measured."""

    def __init__(self, wl, slices):
        import torch
        from oracle import oracle as o
        self.o, self.torch, self.wl, self.slices = o, torch, wl, max(1, int(slices))
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        N, D = wl["n_items"], wl["D"]
        self.n_old = wl["n_old_items"]
        g = np.random.Generator(np.random.PCG64(7))
        self.item_table = (g.standard_normal((self.n_old, D), dtype=np.float32) * np.float32(0.05))
        self.user_table = (g.standard_normal((wl["n_old_users"], D), dtype=np.float32) * np.float32(0.05))
        self.users, self.hu, self.hi = query_batch(wl, 99)
        if wl["embedder"] == "dhe":
            keys = o.keys_to_array(dhe_keys(wl["H"]))
            ws, bs = dhe_weights(wl["H"], wl["hidden"], D, 5)
            tws = [torch.from_numpy(w) for w in ws]
            tbs = [torch.from_numpy(b) for b in bs]

            def mlp(h):          # fp32 torch CPU GEMMs (multi-threaded), like the reference's nn.Sequential on CPU
                x = torch.from_numpy(h.astype(np.float32))
                for l in range(4):
                    x = torch.nn.functional.linear(x, tws[l], tbs[l])
                    x = torch.nn.functional.gelu(x) if l < 3 else torch.sigmoid(x)
                return x.numpy()

            self.embed_items = lambda ids: mlp(o.dhe_hashes(ids, keys))
            self.embed_users = self.embed_items
            self.hashing = "C SipHash loop (the reference loops per id in Python, dh_embedder.py:165-170)"
        else:
            F_, B = wl["F"], wl["B"]
            n_oov = N - self.n_old
            feat = np.empty((N, F_), dtype=np.float32)          # rows < n_old are never hashed
            feat[self.n_old:] = o.l2_normalize(g.standard_normal((n_oov, F_), dtype=np.float32))
            ufeat = o.l2_normalize(g.standard_normal((wl["n_users"], F_), dtype=np.float32))
            planes = g.standard_normal((B, F_), dtype=np.float32)
            W = (g.standard_normal((B, D), dtype=np.float32) * np.float32(0.05))
            chunk = 131072                                       # rows per [n, B] multi-hot block (0.5 GB fp32)

            def emb(fm, ids):
                out = np.empty((len(ids), D), dtype=np.float32)
                for c0 in range(0, len(ids), chunk):
                    out[c0:c0 + chunk] = o.lsh_embed(fm, ids[c0:c0 + chunk], planes, W)
                return out

            self.embed_items = lambda ids: emb(feat, ids)
            self.embed_users = lambda ids: emb(ufeat, ids)
            self.hashing = "torch_hash.py:55-60 restated with numpy GEMMs"
        self.N = N
        self.i = 0
        self.ue = None
        self.best = None

    def step(self):
        """One slice of the item axis (slice 0 also embeds the query users); returns True when a batch completed.
        Slice s holds the items s, s + slices, s + 2 slices, ...: every slice has the table's own mix of in-vocab and OOV
        rows, so every step costs the same and any number of timed steps is an unbiased sample of the batch."""
        o, torch, wl = self.o, self.torch, self.wl
        k = wl["k"]
        S = self.slices
        sl = self.i % S
        if sl == 0:
            self.ue = torch.from_numpy(o.assemble_rows(self.users, wl["n_old_users"], self.user_table, self.embed_users))
            self.best = None
        ids = np.arange(sl, self.N, S)
        ie = o.assemble_rows(ids, self.n_old, self.item_table, self.embed_items)
        s = self.ue @ torch.from_numpy(ie).T
        if sl == 0:
            s[:, 0] = -np.inf                                   # pad item 0
        m = (self.hi % S) == sl
        s[torch.from_numpy(self.hu[m]), torch.from_numpy((self.hi[m] - sl) // S)] = -np.inf
        v, ix = torch.topk(s, min(k, len(ids)), dim=-1)
        ix = ix * S + sl
        if self.best is not None:
            v = torch.cat([self.best[0], v], dim=1)
            ix = torch.cat([self.best[1], ix], dim=1)
            v, sel = torch.topk(v, k, dim=-1)
            ix = torch.gather(ix, 1, sel)
        self.best = (v, ix)
        self.i += 1
        return sl == S - 1

    def run(self, steps, warmup, budget_s=None):
        """Times `steps` slices after `warmup` untimed ones (stops early after `budget_s` seconds, never mid-batch when
        a budget is given).  Returns (queries/s, ms per step, steps timed)."""
        for _ in range(warmup):
            self.step()
        t_all = time.perf_counter()
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            done = self.step()
            times.append(time.perf_counter() - t0)
            if budget_s is not None and done and time.perf_counter() - t_all > budget_s:
                break
        t = float(np.mean(times))
        return self.wl["Q"] / (t * self.slices), t * 1e3, len(times)

    def describe(self, n_steps, ms):
        wl = self.wl
        return (f"oracle port of the reference step on {self.threads} host threads at the full size (N={wl['n_items']}, Q={wl['Q']}): "
                f"{self.slices} strided item-axis slice(s) (items s, s + {self.slices}, ...: same in-vocab / OOV mix) per query batch with a running top-k merge, one slice per step, {n_steps} steps of "
                f"{ms:.0f} ms measured, nothing extrapolated; hashing = {self.hashing}")


def default_slices(wl):
    return 8 if wl["n_items"] >= 5_000_000 else 1


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled every ~5 ms by a thread (NVML) for the whole run; `window(t0, t1)` reports the
    samples that fall inside the timed region (falls back to all samples taken under load if the region was shorter
    than the sampling period)."""

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []          # (time, sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.samples.append((time.time(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def window(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if not self.samples:
            return out
        sel = [x for x in self.samples if t0 <= x[0] <= t1]
        scope = "timed region"
        if len(sel) < 3:                      # region shorter than a few sampling periods: widen to the loaded phase
            sel = [x for x in self.samples if t0 - 0.25 <= x[0] <= t1 + 0.25] or self.samples
            scope = "timed region +-0.25 s"
        nv = self._nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        mask = 0
        for _, _, r in sel:
            mask |= r
        out.update(sm_mhz=float(np.median([x[1] for x in sel])), reasons=sorted(n for n, bit in names.items() if mask & bit),
                   samples=len(sel), scope=scope)
        return out


# --------------------------------------------------------------------------------------- GPU arm
def build_gpu(wl, device, rank):
    import torch
    import oov_b200
    from oov_b200 import ops

    class Config(dict):
        def __getitem__(self, key):
            return dict.get(self, key, None)

    class Dataset:
        def __init__(self, nu, ni, uf, itf):
            self._n = {"user_id": nu, "item_id": ni}
            self.user_num, self.item_num, self._uf, self._if = nu, ni, uf, itf

        def num(self, f):
            return self._n[f]

        def get_user_feature(self):
            return self._uf

        def get_item_feature(self):
            return self._if

    D = wl["D"]
    g = torch.Generator(device="cpu").manual_seed(2020)
    cfg = Config(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device=device, embedding_size=D,
                 add_oov_buckets=True, inductive_embedder=wl["embedder"], user_oov_buckets=wl.get("B", 1000),
                 item_oov_buckets=wl.get("B", 1000), dhe_num_hashes=wl.get("H", 128), table_dtype="bfloat16",
                 oov_normalization_type="global", topk=[10, wl["k"]], gamma=1.0)
    if wl["embedder"] == "dhe":
        # DHE reads no features; tiny placeholders keep the reference constructor contract (dh_embedder.py:91-92)
        uf = oov_b200.Interaction({"user_id": torch.arange(wl["n_users"]), "f0": torch.ones(wl["n_users"], 1)})
        itf = oov_b200.Interaction({"item_id": torch.arange(wl["n_items"]), "f0": torch.ones(wl["n_items"], 1)})
        tmp = tempfile.mkdtemp(prefix=f"oov_bench_r{rank}_")
        os.makedirs(os.path.join(tmp, "hash_keys"))
        with open(os.path.join(tmp, "hash_keys", f"{wl['H']}.hashes"), "w") as f:
            json.dump([k.hex() for k in dhe_keys(wl["H"])], f)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            emb = oov_b200.get_inductive_embedder(cfg, Dataset(wl["n_old_users"], wl["n_old_items"], uf, itf), mode="bench")
        finally:
            os.chdir(cwd)
    else:
        F_ = wl["F"]
        uf = oov_b200.Interaction({"user_id": torch.arange(wl["n_users"]), "f0": torch.randn(wl["n_users"], F_, generator=g)})
        itf = oov_b200.Interaction({"item_id": torch.arange(wl["n_items"]), "f0": torch.randn(wl["n_items"], F_, generator=g)})
        emb = oov_b200.get_inductive_embedder(cfg, Dataset(wl["n_old_users"], wl["n_old_items"], uf, itf), mode=f"bench-{rank}")
    cls = oov_b200.BPR if wl["model"] == "BPR" else oov_b200.DirectAU
    model = cls(cfg, Dataset(wl["n_old_users"], wl["n_old_items"], uf, itf), inductive_embedder=emb).to(device).eval()
    if wl["embedder"] == "dhe":
        # 'trained-looking' hash nets, set AFTER the model is built: the model's constructor re-initialises every Linear of
        # its sub-modules, the embedder's nets included (bpr.py:46 self.apply(xavier_normal_initialization)), which with
        # 24-bit hash inputs saturates every output to exactly 0 or 1
        ws, bs = dhe_weights(wl["H"], wl["hidden"], D, 5)
        with torch.no_grad():
            for net in (emb.user_hash_net, emb.item_hash_net):
                for l, li in enumerate((0, 2, 4, 6)):
                    net[li].weight.copy_(torch.from_numpy(ws[l]))
                    net[li].bias.copy_(torch.from_numpy(bs[l]))
    return cfg, emb, model


def config_json(name, wl, args, world):
    return {"workload": name, "model": wl["model"], "embedder": wl["embedder"], "n_items": wl["n_items"],
            "n_oov_items": wl["n_items"] - wl["n_old_items"], "n_users": wl["n_users"], "embedding_size": wl["D"],
            "Q_per_step": wl["Q"], "k": wl["k"], "max_history": wl["max_hist"],
            "step": "embed Q users + embed all N items + score + mask + top-k (un-amortised, as bpr.py:151-156)",
            "l2": "per-step working set (tables + activations) > 126 MB L2; no explicit flush",
            "launch": "eager (one Python call per kernel)" if args.eager else "whole step replayed from one CUDA graph (GraphedTopK)",
            "parallelism": "single GPU" if world == 1 else f"item rows sharded over {world} ranks, all-gather top-k merge"}


def traffic_of(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel / launch shape, taken from the committed
    `ncu --set full` summaries (profiles/traffic.json names the source file of every entry)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(kernel_key)
        return (float(e["bytes"]), e["source"]) if e else (None, None)
    except Exception:
        return None, None


def run_reference(args, name, wl):
    """`--impl reference`: the CPU arm alone (rank 0), same metric / unit / config keys as the GPU arm."""
    slices = args.cpu_slices or default_slices(wl)
    ref = CpuReference(wl, slices)
    qps, ms, n = ref.run(args.steps, args.warmup)
    desc = ref.describe(n, ms)
    return {"impl": "reference", "metric": "full_sort_topk_queries_per_s", "value": qps, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_json(name, wl, args, max(args.gpus, 1)),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": ref.threads, "kind": "port", "sample": desc,
                             "steps_per_query_batch": slices},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_gpu(args, name, wl, rank, world, local_rank, with_cpu_baseline):
    """One workload on this rank's GPU (all ranks call it); returns the result dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    import oov_b200
    from oov_b200 import ops, sharded
    device = f"cuda:{local_rank}"
    cfg_json = config_json(name, wl, args, world)
    cfg, emb, model = build_gpu(wl, device, rank)
    Q, k, N = wl["Q"], wl["k"], wl["n_items"]
    n_batches = 8
    batches = [query_batch(wl, 100 + b) for b in range(n_batches)]
    dev_batches = []
    for users, hu, hi in batches:
        u = torch.from_numpy(users).to(device)
        dhu, dhi = torch.from_numpy(hu).to(device), torch.from_numpy(hi).to(device)
        csr = ops.pairs_to_csr(dhu, dhi, Q)
        dev_batches.append((u, csr, dhu, dhi))
    pin = [(torch.from_numpy(u).pin_memory(), torch.from_numpy(hu).pin_memory(), torch.from_numpy(hi).pin_memory())
           for u, hu, hi in batches]
    out_s_host = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((Q, k), dtype=torch.int64).pin_memory()

    sr = sharded.ShardedRetrieval(model, N) if world > 1 else None
    # the public serving call: the whole step (history CSR build, user embed, item-table assembly, scoring, masks,
    # top-k, candidate all-gather + merge when sharded) captured once in a CUDA graph and replayed per batch
    gstep = None if args.eager else oov_b200.GraphedTopK(model, Q, k, N, Q * wl["max_hist"], sharded=sr)

    def step_resident(b):
        u, csr, dhu, dhi = dev_batches[b % n_batches]
        if gstep is not None:
            return gstep(u, dhu, dhi)
        if sr is None:
            return model.full_sort_topk(u, k, n_total_items=N, hist_csr=csr)
        user_e = model._assemble("user", u, out_dtype=model.table_dtype)
        sr.build_shard()
        return sr.topk(user_e, k, hist=csr)

    def step_e2e(b):
        hu_, hhu, hhi = pin[b % n_batches]
        if gstep is not None:
            s, i = gstep(hu_, hhu, hhi)                       # pinned host -> static device buffers, then the graph
        else:
            u = hu_.to(device, non_blocking=True)
            hu = hhu.to(device, non_blocking=True)
            hi = hhi.to(device, non_blocking=True)
            if sr is None:
                s, i = model.full_sort_topk({"user_id": u}, k, n_total_items=N, history_index=(hu, hi))
            else:
                csr = ops.pairs_to_csr(hu, hi, Q)
                user_e = model._assemble("user", u, out_dtype=model.table_dtype)
                sr.build_shard()
                s, i = sr.topk(user_e, k, hist=csr)
        out_s_host.copy_(s, non_blocking=True)
        out_i_host.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller consumes the result every step
        return hu_.numel() * 8 + hhu.numel() * 8 + hhi.numel() * 8, s.numel() * 4 + i.numel() * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in range(steps):
            fn(b)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for b in range(max(args.warmup, 3)):
        step_resident(b)
        step_e2e(b)
    l0 = ops.launch_count()
    t_w0 = time.time()
    ms_total = timed(step_resident, args.steps)
    t_w1 = time.time()
    launches = ops.launch_count() - l0 if gstep is None else gstep.launches_per_replay * args.steps
    clocks = sampler.window(t_w0, t_w1) if sampler else None
    h2d = d2h = 0

    def e2e_fn(b):
        nonlocal h2d, d2h
        h2d, d2h = step_e2e(b)
    ms_e2e = timed(e2e_fn, args.steps)
    lt = torch.tensor([launches], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(lt)
    value = Q * args.steps / (ms_total * 1e-3)
    e2e_value = Q * args.steps / (ms_e2e * 1e-3)

    # ---- per-stage split + dominant kernel, timed with CUDA events on the launching stream (rank 0, after the run)
    stages, roofline = {}, None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass

        def ev_time(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        u, csr = dev_batches[0][0], dev_batches[0][1]
        table = model.build_item_table(N)
        user_e = model._assemble("user", u, out_dtype=model.table_dtype)
        stages["user_embed_ms"] = ev_time(lambda: model._assemble("user", u, out_dtype=model.table_dtype))
        stages["item_table_ms"] = ev_time(lambda: model.build_item_table(N), reps=3)
        stages["score_topk_ms"] = ev_time(lambda: ops.fullsort_topk(user_e, table, k, hist=csr))
        n_oov = N - wl["n_old_items"]
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        src = "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1590 (of fallback)"
        cands = []          # (per-step ms, roofline dict)

        def tensor_entry(name_, key, t_ms, flops, launches_per_step=1, note=None):
            ach = flops / (t_ms * 1e-3) / 1e12
            tr, tr_src = traffic_of(key)
            d = {"kernel": name_, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                 "traffic": tr, "launch_ms": t_ms, "flops_per_launch": flops, "launches_per_step": launches_per_step,
                 "peak_source": src}
            if tr_src:
                d["traffic_source"] = tr_src
            if note:
                d["note"] = note
            cands.append((t_ms * launches_per_step, d))

        # fused score + mask + top-k: algorithmic flops 2 Q N D
        tensor_entry("tc_score_topk_kernel<1> + score_threshold_kernel + tc_score_topk_kernel<0> + merge_keys_kernel "
                     "(tcgen05 scoring fused with masks and top-k; sampled pre-pass threshold)", f"score_topk:{name}",
                     stages["score_topk_ms"], 2.0 * Q * N * wl["D"],
                     note="algorithmic flops 2 Q N D over the time of all four launches; the accumulator hand-over "
                          "(D = 64: four MMAs per 128 x 128 tile) and the epilogue bound it, not the MMA rate")
        if wl["embedder"] == "dhe":
            ids_oov = torch.arange(wl["n_old_items"], N, device=device)
            keys_dev = emb._keys_dev
            stages["dhe_hash_u32_ms"] = ev_time(lambda: ops.dhe_hash(ids_oov, keys_dev), reps=3)   # generic oov_dhe_hash; the step uses the byte-plane kernel inside oov_dhe_embed
            M = min(n_oov, 1 << 18)
            A = torch.randn(M, wl["hidden"], device=device).to(torch.bfloat16)
            Wt = torch.randn(wl["hidden"], wl["hidden"], device=device).to(torch.bfloat16)
            bias = torch.zeros(wl["hidden"], device=device)
            t_ms = ev_time(lambda: ops.tc_linear(A, Wt, bias, act="gelu", out_dtype=torch.bfloat16), reps=10)
            stages["dhe_hidden_layer_ms_per_262144_rows"] = t_ms
            tensor_entry("tc_linear2_kernel<GELU> (DHE hidden layer 512x512, tcgen05 cta_group::2 CTA pairs)",
                         f"tc_linear2_gelu:M{M}:N{wl['hidden']}:K{wl['hidden']}", t_ms,
                         2.0 * M * wl["hidden"] * wl["hidden"], launches_per_step=2 * max(1, -(-n_oov // M)))
        else:
            ids_oov = torch.arange(wl["n_old_items"], N, device=device)
            feat_i = emb.item_feature_mat
            planes_i = emb.item_lsh.uniform_planes[0].data
            Wb = model.item_oov_buckets.weight.data
            out_b = torch.empty((n_oov, wl["D"]), dtype=torch.bfloat16, device=device)
            t_ms = ev_time(lambda: ops.lsh_embed(feat_i, planes_i, Wb, ids_oov, out=out_b, n_old=0), reps=3)
            stages["lsh_embed_oov_ms"] = t_ms
            tensor_entry("tc_lsh_embed_kernel (sign-projection GEMM + bucket-mean GEMM, tcgen05, operands in TMEM)",
                         f"tc_lsh_embed:n{n_oov}:F{wl['F']}:B{wl['B']}:D{wl['D']}", t_ms,
                         2.0 * n_oov * wl["B"] * (wl["F"] + wl["D"]),
                         note="algorithmic flops = 2 B (F + D) per OOV id (SURVEY 8d); the kernel issues 3 F + D wide fp16 MMAs (exact-sign split)")
        # throughput when the item table is embedded once per weight version and reused by every query
        # batch (only the query side and the scoring run per step) — derived from the stage timings above
        stages["amortised_queries_per_s_item_table_reused"] = Q / ((stages["user_embed_ms"] + stages["score_topk_ms"]) * 1e-3)
        roofline = max(cands, key=lambda c: c[0])[1]
        roofline["others"] = [c[1] for c in cands if c[1] is not roofline]
        del table, user_e

    cpu_base = None
    if rank == 0 and with_cpu_baseline:
        # bounded: at most one full query batch beyond ~15 s of CPU work, at the full workload size (nothing extrapolated)
        slices = args.cpu_slices or default_slices(wl)
        ref = CpuReference(wl, slices)
        qps, ms_c, n_c = ref.run(1000 * slices, 0, budget_s=15.0)
        cpu_base = {"value": qps, "unit": "queries/s", "cores": ref.threads, "kind": "port", "sample": ref.describe(n_c, ms_c),
                    "steps_per_query_batch": slices, "ms_per_step": ms_c}
        del ref

    result = None
    if rank == 0:
        result = {
            "metric": "full_sort_topk_queries_per_s", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": cfg_json, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(lt.item()), "roofline": roofline, "cpu_baseline": cpu_base, "stages": stages}
    if sampler:
        sampler.stop()
    # free this workload's device memory before the next one (a captured graph holds its pool until it is dropped)
    gstep = None
    del dev_batches, model, emb, sr
    import gc
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return result


DCNV2_WL = dict(rows=65536, fields=26, D=16, mlp=[768, 768], cross_layers=3, n_users=100_000, n_old_users=50_000,
                n_items=100_000, n_old_items=50_000, other_vocab=100_000, F=32, B=1000)


def run_dcnv2(args, device):
    """BASELINE.json configs[2]: DCNV2 ranking with slsh OOV buckets on Criteo-shaped synthetic rows —
    26 token fields (user, item, 24 more), eval_batch_size 65536, embedding_size 16, cross 3 x 416x416, MLP [768, 768].
    A step = token gather + slsh OOV overwrite (bf16) + cross network + MLP + predict + sigmoid for one batch of rows.
    Reported under `workloads.dcnv2_criteo` as rows/s (N = 1 only: the path has no exchange step; replicas only)."""
    import torch
    import oov_b200
    from oov_b200 import ops
    wl = DCNV2_WL
    Bn, fields, D = wl["rows"], wl["fields"], wl["D"]

    class Config(dict):
        def __getitem__(self, key):
            return dict.get(self, key, None)

    class Dataset:
        def __init__(self, uf, itf):
            self._uf, self._if = uf, itf
            self.user_num, self.item_num = wl["n_old_users"], wl["n_old_items"]

        def num(self, f):
            return {"user_id": self.user_num, "item_id": self.item_num}[f]

        def get_user_feature(self):
            return self._uf

        def get_item_feature(self):
            return self._if

    g = torch.Generator(device="cpu").manual_seed(2021)
    cfg = Config(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", device=device, embedding_size=D, add_oov_buckets=True,
                 inductive_embedder="slsh", user_oov_buckets=wl["B"], item_oov_buckets=wl["B"], oov_normalization_type="global",
                 structure="stacked", cross_layer_num=wl["cross_layers"], mlp_hidden_size=wl["mlp"], dropout_prob=0.2, mixed=False)
    uf = oov_b200.Interaction({"user_id": torch.arange(wl["n_users"]), "f0": torch.randn(wl["n_users"], wl["F"], generator=g)})
    itf = oov_b200.Interaction({"item_id": torch.arange(wl["n_items"]), "f0": torch.randn(wl["n_items"], wl["F"], generator=g)})
    emb = oov_b200.get_inductive_embedder(cfg, Dataset(uf, itf), mode="bench-dcnv2")
    dims = [wl["n_old_users"], wl["n_old_items"]] + [wl["other_vocab"]] * (fields - 2)
    model = oov_b200.DCNV2(cfg, dims, inductive_embedder=emb).to(device).eval()
    with torch.no_grad():                      # keep the cross activations O(1): xavier on a square matrix already is
        for w in model.cross_layer_w:
            w.mul_(1.0 / 416 ** 0.5)
    model.pack_tower()
    n_batches = 4
    host = []
    for b in range(n_batches):
        gb = torch.Generator(device="cpu").manual_seed(300 + b)
        t = torch.stack([torch.randint(0, wl["n_users"], (Bn,), generator=gb), torch.randint(0, wl["n_items"], (Bn,), generator=gb)] +
                        [torch.randint(0, wl["other_vocab"], (Bn,), generator=gb) for _ in range(fields - 2)], dim=1)
        host.append(t.pin_memory())
    dev = [t.to(device) for t in host]
    out_host = torch.empty((Bn,), dtype=torch.float32).pin_memory()

    gstep = None if args.eager else oov_b200.GraphedRanker(model, Bn, fields)

    def step_resident(b):
        return (gstep or model)(dev[b % n_batches])

    def step_e2e(b):
        if gstep is not None:
            out = gstep(host[b % n_batches])                  # pinned host -> the graph's static token buffer
        else:
            out = model(host[b % n_batches].to(device, non_blocking=True))
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in range(steps):
            fn(b)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(int(str(device).split(":")[-1]))
    sampler.start()
    for b in range(max(args.warmup, 3)):
        step_resident(b)
        step_e2e(b)
    l0 = ops.launch_count()
    t0 = time.time()
    ms = timed(step_resident, args.steps)
    t1 = time.time()
    launches = ops.launch_count() - l0 if gstep is None else gstep.launches_per_replay * args.steps
    clocks = sampler.window(t0, t1)
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop()

    def ev_time(fn, reps=10):
        """GPU time of `fn` replayed from a CUDA graph (launched from Python the stages are host-bound)."""
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            keep = fn()
        torch.cuda.synchronize()
        gr.replay()
        ms_ = timed(lambda b: gr.replay(), reps) / reps
        del gr, keep
        return ms_

    x0 = model.embed_token_fields(dev[0], out_dtype=torch.bfloat16).reshape(Bn, -1)
    cross = model.cross_network(x0)
    stages = {"token_gather_oov_ms": ev_time(lambda: model.embed_token_fields(dev[0], out_dtype=torch.bfloat16)),
              "cross_network_ms": ev_time(lambda: model.cross_network(x0)),
              "mlp_ms": ev_time(lambda: model._mlp(cross)),
              "tower_ms": ev_time(lambda: model.tower(x0))}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    inf = fields * D
    flops = 2.0 * Bn * (wl["cross_layers"] * inf * inf + inf * wl["mlp"][0] + wl["mlp"][0] * wl["mlp"][1] + wl["mlp"][1])
    ach = flops / (stages["tower_ms"] * 1e-3) / 1e12
    roofline = {"kernel": "DCNV2 tower: 3 x (tc_linear 416x416 + cross_update) + tc_linear<ReLU> 416->768->768 + predict (tcgen05)",
                "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                "launch_ms": stages["tower_ms"], "flops_per_launch": flops,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops" if peaks else "fallback 1590",
                "note": "algorithmic flops 2 B (3 in^2 + in h1 + h1 h2 + h2) over the time of the whole tower (9 launches); "
                        "the activations (54 - 100 MB per layer) stream through L2 / HBM between launches"}
    res = {"metric": "dcnv2_eval_rows_per_s", "value": Bn * args.steps / (ms * 1e-3), "unit": "rows/s", "ms_per_step": ms / args.steps,
           "e2e": {"value": Bn * args.steps / (ms_e2e * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": Bn * fields * 8,
                   "d2h_bytes_per_step": Bn * 4, "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None, "stages": stages,
           "config": {"workload": "dcnv2_criteo", "model": "DCNV2 (stacked, mixed off, eval)", "embedder": "slsh", "rows_per_step": Bn,
                      "token_fields": fields, "embedding_size": D, "cross_layers": wl["cross_layers"], "mlp_hidden_size": wl["mlp"],
                      "oov_buckets": wl["B"], "oov_share_user_item": 0.5, "l2": "activations per layer 54-100 MB + 4 rotating batches; no explicit flush",
                      "launch": "eager" if gstep is None else "whole forward replayed from one CUDA graph (GraphedRanker)",
                      "parallelism": "single GPU (replicas only: no exchange step)"},
           "clocks": clocks, "steps": args.steps, "warmup": max(args.warmup, 3)}
    gstep = None
    del model, emb, dev
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return res


XDEEPFM_WL = dict(rows=65536, fields=26, D=10, mlp=[128, 128, 128], cin=[100, 100, 100], n_users=100_000, n_old_users=50_000,
                  n_items=100_000, n_old_items=50_000, other_vocab=100_000)


def run_xdeepfm(args, device):
    """BASELINE.json configs[3]: xDeepFM ranking with the `mean` OOV embedder on the same
    Criteo-shaped synthetic rows — 26 token fields, batch 65536, embedding_size 10, CIN [100, 100, 100] (direct off), MLP
    [128, 128, 128], first-order linear with its own `mean` embedder.  A step = token gather + OOV overwrite (bf16) +
    first-order sum + CIN + MLP for one batch.  Reported under `workloads.xdeepfm_criteo` as rows/s (N = 1 only)."""
    import torch
    import oov_b200
    from oov_b200 import ops
    wl = XDEEPFM_WL
    Bn, fields, D = wl["rows"], wl["fields"], wl["D"]

    class Config(dict):
        def __getitem__(self, key):
            return dict.get(self, key, None)

    class Dataset:
        def __init__(self, uf, itf):
            self._uf, self._if = uf, itf
            self.user_num, self.item_num = wl["n_old_users"], wl["n_old_items"]

        def num(self, f):
            return {"user_id": self.user_num, "item_id": self.item_num}[f]

        def get_user_feature(self):
            return self._uf

        def get_item_feature(self):
            return self._if

    cfg = Config(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", device=device, embedding_size=D, add_oov_buckets=True,
                 inductive_embedder="mean", user_oov_buckets=10, item_oov_buckets=10, mlp_hidden_size=wl["mlp"],
                 cin_layer_size=wl["cin"], direct=False, dropout_prob=0.2)
    uf = oov_b200.Interaction({"user_id": torch.arange(wl["n_users"]), "f0": torch.ones(wl["n_users"], 1)})
    itf = oov_b200.Interaction({"item_id": torch.arange(wl["n_items"]), "f0": torch.ones(wl["n_items"], 1)})
    ds = Dataset(uf, itf)
    emb = oov_b200.get_inductive_embedder(cfg, ds, mode="bench-xdeepfm")
    emb1 = oov_b200.get_inductive_embedder(cfg, ds, mode="bench-xdeepfm", embedding_size=1, first_order=True)
    dims = [wl["n_old_users"], wl["n_old_items"]] + [wl["other_vocab"]] * (fields - 2)
    torch.manual_seed(2022)
    model = oov_b200.xDeepFM(cfg, dims, inductive_embedder=emb, first_order_embedder=emb1).to(device).eval()
    with torch.no_grad():                      # embeddings of O(0.3) keep the three chained Hadamard layers in bf16 range
        model.token_embedding_table.embedding.weight.mul_(0.3 / model.token_embedding_table.embedding.weight.std())
    model.pack_tower()
    n_batches = 4
    host = []
    for b in range(n_batches):
        gb = torch.Generator(device="cpu").manual_seed(400 + b)
        t = torch.stack([torch.randint(0, wl["n_users"], (Bn,), generator=gb), torch.randint(0, wl["n_items"], (Bn,), generator=gb)] +
                        [torch.randint(0, wl["other_vocab"], (Bn,), generator=gb) for _ in range(fields - 2)], dim=1)
        host.append(t.pin_memory())
    dev = [t.to(device) for t in host]
    out_host = torch.empty((Bn,), dtype=torch.float32).pin_memory()
    class Predict(torch.nn.Module):            # the graph replays `predict` (sigmoid of the logits), like the evaluator calls it
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, tok):
            return self.m.predict(tok)

    pred = Predict(model)
    gstep = None if args.eager else oov_b200.GraphedRanker(pred, Bn, fields)

    def step_resident(b):
        return (gstep or pred)(dev[b % n_batches])

    def step_e2e(b):
        if gstep is not None:
            out = gstep(host[b % n_batches])
        else:
            out = pred(host[b % n_batches].to(device, non_blocking=True))
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in range(steps):
            fn(b)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(int(str(device).split(":")[-1]))
    sampler.start()
    for b in range(max(args.warmup, 3)):
        step_resident(b)
        step_e2e(b)
    l0 = ops.launch_count()
    t0 = time.time()
    ms = timed(step_resident, args.steps)
    t1 = time.time()
    launches = ops.launch_count() - l0 if gstep is None else gstep.launches_per_replay * args.steps
    clocks = sampler.window(t0, t1)
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop()

    def ev_time(fn, reps=5):
        for _ in range(2):
            fn()
        return timed(lambda b: fn(), reps) / reps

    x0 = model.embed_token_fields(dev[0], out_dtype=torch.bfloat16)
    stages = {"token_gather_oov_ms": ev_time(lambda: model.embed_token_fields(dev[0], out_dtype=torch.bfloat16)),
              "first_order_ms": ev_time(lambda: model.first_order_linear(dev[0])),
              "cin_ms": ev_time(lambda: model.compressed_interaction_network(x0)),
              "mlp_ms": ev_time(lambda: model.deep(x0.reshape(Bn, -1)))}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    hs = [fields] + [c // 2 for c in wl["cin"][:-1]]
    flops = 2.0 * Bn * D * sum(h * fields * c for h, c in zip(hs, wl["cin"]))
    zbytes = 2.0 * Bn * D * sum(2 * ((h * fields + 7) // 8 * 8) for h in hs)          # z written once and read once, bf16
    ach = flops / (stages["cin_ms"] * 1e-3) / 1e12
    fused = bool(model.fused_cin and model._cin_fusable(fields))
    roofline = {"kernel": "xDeepFM CIN: 3 x tc_cin_layer_kernel (outer-product operand generated in shared memory, tcgen05, pooled epilogue)" if fused
                else "xDeepFM CIN: 3 x (cin_outer + tc_linear<ReLU> + cin_pool_dot), 8 chunks of 8192 rows", "bound": "tensor",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                "launch_ms": stages["cin_ms"], "flops_per_launch": flops,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops" if peaks else "fallback 1590",
                "note": "algorithmic flops 2 B D sum_k H_{k-1} M H_k over the time of the whole CIN (three launches + the d-major "
                        "transpose of the embeddings); the issued MMA work is 1.6x that (N padded 100 -> 128, field pitch 26 -> 32). "
                        + ("The operand z never reaches HBM; the kernel is bound by the per-k-block hand-over chain (DESIGN 4.8)." if fused else
                           f"The outer-product operand z ({zbytes / 1e9:.1f} GB written + read per batch) goes through HBM.")}
    res = {"metric": "xdeepfm_eval_rows_per_s", "value": Bn * args.steps / (ms * 1e-3), "unit": "rows/s", "ms_per_step": ms / args.steps,
           "e2e": {"value": Bn * args.steps / (ms_e2e * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": Bn * fields * 8,
                   "d2h_bytes_per_step": Bn * 4, "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None, "stages": stages,
           "config": {"workload": "xdeepfm_criteo", "model": "xDeepFM (direct off, eval)", "embedder": "mean (+ first-order mean)", "rows_per_step": Bn,
                      "token_fields": fields, "embedding_size": D, "cin_layer_size": wl["cin"], "mlp_hidden_size": wl["mlp"],
                      "l2": "z operands 0.2 - 0.4 GB per chunk + 4 rotating batches; no explicit flush",
                      "launch": "eager" if gstep is None else "whole forward replayed from one CUDA graph (GraphedRanker)",
                      "parallelism": "single GPU (replicas only: no exchange step)"},
           "clocks": clocks, "steps": args.steps, "warmup": max(args.warmup, 3)}
    gstep = None
    del model, emb, emb1, dev
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return res


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    names = [args.workload] + ([] if args.single or args.second == args.workload else [args.second])

    def workload(name):
        wl = dict(WORKLOADS[name])
        if args.Q:
            wl["Q"] = args.Q
        return wl

    if args.impl == "reference":
        if rank != 0:
            return
        print(json.dumps(run_reference(args, names[0], workload(names[0]))))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    import oov_b200
    oov_b200._lib.check(oov_b200._lib.load().oov_check_device(local_rank))

    results = [run_gpu(args, n, workload(n), rank, world, local_rank, with_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
               for n in names]
    dcn = run_dcnv2(args, f"cuda:{local_rank}") if (world == 1 and args.dcnv2 and not args.single) else None
    xdf = run_xdeepfm(args, f"cuda:{local_rank}") if (world == 1 and args.dcnv2 and not args.single) else None
    if rank == 0:
        line = results[0]
        if len(results) > 1:
            line.setdefault("workloads", {}).update({n: {key: r[key] for key in ("value", "unit", "ms_per_step", "e2e", "gpu_launches", "roofline",
                                                           "cpu_baseline", "stages", "config", "clocks", "steps", "warmup")}
                                 for n, r in zip(names[1:], results[1:])})
        if dcn is not None:
            line.setdefault("workloads", {})["dcnv2_criteo"] = dcn
        if xdf is not None:
            line.setdefault("workloads", {})["xdeepfm_criteo"] = xdf
        print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        # A captured graph that contains the NCCL all-gather keeps the communicator busy at teardown
        # (destroy_process_group did not return on the 2-GPU box): quiesce, leave without the interpreter's teardown.
        # Every rank has finished its work and rank 0 has printed by now.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
