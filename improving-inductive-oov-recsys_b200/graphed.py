"""Static-shape full-sort top-k step captured in ONE CUDA graph.

A retrieval step is ~25-60 kernel launches (user embed, item-table assembly, pre-pass / threshold / scoring /
merge per segment, the candidate all-gather when sharded).  On a small shard (1M items over 8 GPUs) the kernels are
tens of microseconds each and the host enqueue time (~0.45 ms) is close to the GPU time, so the whole step is
captured once and replayed: the host cost per step drops to the input copies plus one `cudaGraphLaunch`.

No reference counterpart (the reference evaluates eagerly, trainer.py:526-545); the inputs and outputs are the same
as `model.full_sort_topk`: user ids [Q] and the dataloader's history_index pairs (general_dataloader.py:270-292).
History pairs are padded to `max_pairs` with row = Q (dropped by the CSR build), so every shape is static.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class GraphedTopK:
    """`scores, ids = g(users, hist_rows, hist_cols)` — replays the captured step.

    model         BPR / DirectAU (model/general.py)
    Q, k          static query batch size and list length
    n_total_items N of `get_item_embedding(arange(N))`
    max_pairs     capacity of the history pair buffers
    sharded       optional `ShardedRetrieval` (row-sharded step with the NCCL all-gather inside the graph)
    """

    def __init__(self, model, Q: int, k: int, n_total_items: int, max_pairs: int, sharded=None, warmup: int = 3):
        self.model, self.Q, self.k, self.N, self.max_pairs, self.sr = model, Q, k, n_total_items, max(int(max_pairs), 1), sharded
        dev = model.device
        self.users = torch.zeros(Q, dtype=torch.int64, device=dev)
        self.hist_rows = torch.full((self.max_pairs,), Q, dtype=torch.int64, device=dev)
        self.hist_cols = torch.zeros(self.max_pairs, dtype=torch.int64, device=dev)
        self.scores: Optional[torch.Tensor] = None
        self.ids: Optional[torch.Tensor] = None
        self.launches_per_replay = 0
        self._side = torch.cuda.Stream(device=dev)                 # query-side branch of the step
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):                        # allocator, workspaces, function attributes, NCCL
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count()
        # capture on the SAME stream the warm-up ran on: the per-(device, stream) kernel workspaces (ops._workspace) the
        # graph bakes in are then exactly the ones the warm-up sized, owned by this object's two streams
        with torch.cuda.graph(self.graph, stream=side):
            self.scores, self.ids = self._step()
        self.launches_per_replay = ops.launch_count() - l0
        torch.cuda.synchronize(dev)
        # the graph holds raw addresses of those workspaces: keep them alive even if the cache later replaces them with
        # bigger buffers for some other caller of the same stream
        streams = {side.cuda_stream, self._side.cuda_stream}
        self._keep_alive = [buf for key, buf in ops._ws_cache.items() if key[1] in streams]
        self._capture_stream = side

    def _step(self):
        """One step on the current stream, with the query side (history CSR + user embed: a dozen small kernels that
        fill a few SMs) forked onto a second stream so that it runs under the item-table assembly instead of in front
        of it; the streams join before the scoring launch.  Captured as two branches of the graph."""
        m = self.model
        cur = torch.cuda.current_stream(m.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            if self.sr is not None and self.sr.fused:
                csr = self.sr.local_history_csr((self.hist_rows, self.hist_cols), self.Q)
            else:
                csr = ops.pairs_to_csr(self.hist_rows, self.hist_cols, self.Q)      # padding rows (= Q) are dropped
            user_e = m._assemble("user", self.users, out_dtype=m.table_dtype)
        if self.sr is None:
            split = min(m.n_items, self.N)
            if m.inductive_embedder is not None and m.inductive_mapper is None and 0 < split < self.N:
                # the in-vocab half of the table (a memory-bound gather-cast) joins the query-side branch, the OOV half
                # (SipHash + MLP / LSH GEMMs) is the main branch
                table = torch.empty((self.N, m.embedding_size), dtype=m.table_dtype, device=m.device)
                if not m.build_item_rows_fused((0, split), table[:split], (split, self.N), table[split:]):
                    with torch.cuda.stream(self._side):
                        m.build_item_table(self.N, row_range=(0, split), out=table[:split])
                    m.build_item_table(self.N, row_range=(split, self.N), out=table[split:])
            else:
                table = m.build_item_table(self.N)
            cur.wait_stream(self._side)
            s, i = ops.fullsort_topk(user_e, table, self.k, mask_pad=True, hist=csr)
        else:
            if self.sr.fused:
                self.sr.build_shard(iv_stream=self._side)          # in-vocab slice on the query-side branch
            else:
                self.sr.build_shard()
            cur.wait_stream(self._side)
            if self.sr.fused:
                s, i = self.sr.topk(user_e, self.k, local_csr=csr)
            else:
                s, i = self.sr.topk(user_e, self.k, hist=csr)
        self.hist_rows.fill_(self.Q)                               # padding for the next call's shorter pair list
        return s, i

    def load(self, users: torch.Tensor, hist_rows: Optional[torch.Tensor], hist_cols: Optional[torch.Tensor]) -> None:
        """Copy one batch into the static buffers (host pinned or device tensors; async on the current stream)."""
        if users.shape[0] != self.Q:
            raise ValueError(f"graph was captured for Q={self.Q}, got {users.shape[0]} users")
        self.users.copy_(users, non_blocking=True)
        n = 0 if hist_rows is None else int(hist_rows.shape[0])
        if n > self.max_pairs:
            raise ValueError(f"{n} history pairs exceed the captured capacity {self.max_pairs}")
        if n:
            self.hist_rows[:n].copy_(hist_rows, non_blocking=True)
            self.hist_cols[:n].copy_(hist_cols, non_blocking=True)

    def __call__(self, users, hist_rows=None, hist_cols=None):
        self.load(users, hist_rows, hist_cols)
        self.graph.replay()
        return self.scores, self.ids


class GraphedRanker:
    """`probs = g(token_fields)` — a context ranking model's forward (token gather + OOV overwrite + dense tower, e.g.
    `model.context.DCNV2`) for a static batch of `rows` x `fields` token ids, captured once and replayed.  The tower is a
    dozen launches of 20-60 us each: launched one by one from Python the host enqueue time exceeds the GPU time.
    A shorter batch is padded with token 0 rows and the result sliced."""

    def __init__(self, model, rows: int, fields: int, warmup: int = 3):
        self.model, self.rows, self.fields = model, int(rows), int(fields)
        dev = next(model.parameters()).device
        self.tokens = torch.zeros((self.rows, self.fields), dtype=torch.int64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                model(self.tokens)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count()
        with torch.cuda.graph(self.graph, stream=side):
            self.out = model(self.tokens)
        self.launches_per_replay = ops.launch_count() - l0
        torch.cuda.synchronize(dev)
        self._keep_alive = [buf for key, buf in ops._ws_cache.items() if key[1] == side.cuda_stream]
        self._capture_stream = side

    def __call__(self, token_fields: torch.Tensor) -> torch.Tensor:
        n = int(token_fields.shape[0])
        if n > self.rows or token_fields.shape[1] != self.fields:
            raise ValueError(f"graph was captured for [{self.rows}, {self.fields}] token ids, got {tuple(token_fields.shape)}")
        self.tokens[:n].copy_(token_fields, non_blocking=True)
        if n < self.rows:
            self.tokens[n:].zero_()
        self.graph.replay()
        return self.out[:n]
