"""Row-sharded full-sort retrieval over N GPUs — no reference counterpart: the
reference's inductive path is single-process (its driver never sets local_rank, configurator.py:487-494).

One process per GPU (`torch.distributed`, NCCL over NVLink).  Items are independent rows, so the item
side is row-sharded and the small state (planes, OOV buckets, DHE nets/keys, user tables, the query
batch) is replicated.  Each rank embeds its own rows, runs the fused score+mask+top-k on its shard and
contributes `[S, Q, k]` (score, global id) candidates; ONE all-gather moves 16·S·Q·k bytes per rank and
every rank merges G·S·k -> k with the deterministic (score desc, id asc) rule, so all ranks hold the
same answer.  There is no other data-path collective.

Load balance: new (OOV) ids are appended after the in-vocab ids, and OOV rows are the expensive ones to
embed, so each rank owns one slice of the in-vocab range AND one slice of the OOV range (S = 2 segments)
instead of one contiguous slice of [0, N).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import INT64_MAX


def split_range(lo: int, hi: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-equal split of [lo, hi) — the first (hi-lo) % world ranks get one extra row."""
    n = max(hi - lo, 0)
    base, rem = divmod(n, world)
    start = lo + rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_segments(n_old: int, n_total: int, rank: int, world: int, balanced: bool = True) -> List[Tuple[int, int]]:
    """Item-id segments owned by `rank`.  balanced: [slice of in-vocab ids, slice of OOV ids]."""
    if not balanced:
        return [split_range(0, n_total, rank, world)]
    return [split_range(0, min(n_old, n_total), rank, world), split_range(min(n_old, n_total), n_total, rank, world)]


def pack_candidates(scores: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(fp32 [.., k], int64 [.., k]) -> int64 [.., k, 2]: one buffer, one all-gather."""
    bits = scores.contiguous().view(torch.int32).to(torch.int64)
    return torch.stack([bits, idx], dim=-1).contiguous()


def unpack_candidates(buf: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    scores = buf[..., 0].to(torch.int32).contiguous().view(torch.float32)
    return scores, buf[..., 1].contiguous()


class ShardedRetrieval:
    """Row-sharded `full_sort_topk`.  `model` is a BPR / DirectAU from `model/general.py`.

    `local_topk_fn(user_e, table, k, item_id_offset, seg, hist) -> (scores, ids)` and
    `merge_fn(cand_scores [L,Q,k], cand_idx [L,Q,k]) -> (scores, ids)` default to the CUDA kernels;
    the CPU (gloo) tests inject oracle stand-ins to exercise the exchange logic without a GPU.
    """

    def __init__(self, model, n_total_items: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                 group=None, balanced: bool = True, local_topk_fn: Optional[Callable] = None,
                 merge_fn: Optional[Callable] = None, build_table_fn: Optional[Callable] = None):
        self.model = model
        self.group = group
        inited = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if inited else 0)
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if inited else 1)
        self.n_total = n_total_items
        self.segments = [s for s in shard_segments(model.n_items, n_total_items, self.rank, self.world, balanced)]
        # With the stock kernels the two segments live in ONE table [in-vocab slice | OOV slice] and are scored by ONE
        # fused launch in local-row space (history columns are rewritten to local rows by the CSR build, the winning
        # rows are mapped back to global ids afterwards): half the launches of the per-segment path.
        self.fused = local_topk_fn is None and build_table_fn is None and len(self.segments) == 2
        self.table: Optional[torch.Tensor] = None
        if local_topk_fn is None or merge_fn is None:
            from . import ops
            local_topk_fn = local_topk_fn or (lambda ue, tab, k, off, seg, hist: ops.fullsort_topk(
                ue, tab, k, item_id_offset=off, mask_pad=True, seg=seg, hist=hist))
            merge_fn = merge_fn or ops.topk_merge
        self._local_topk, self._merge = local_topk_fn, merge_fn
        self._build = build_table_fn or (lambda lo, hi: model.build_item_table(n_total_items, row_range=(lo, hi)))
        self.tables: Optional[List[torch.Tensor]] = None

    def build_shard(self, iv_stream: Optional[torch.cuda.Stream] = None) -> List[torch.Tensor]:
        """Embed this rank's rows (in-vocab gather + OOV embed); no communication.  `iv_stream`: run the in-vocab
        segment (a memory-bound gather) on that stream, concurrently with the OOV segment; the caller joins the streams."""
        if self.fused:
            m = self.model
            n0 = self.segments[0][1] - self.segments[0][0]
            n1 = self.segments[1][1] - self.segments[1][0]
            self.table = torch.empty((n0 + n1, m.embedding_size), dtype=m.table_dtype, device=m.device)
            self.tables = [self.table[:n0], self.table[n0:]]
            (a, b), (c, d) = self.segments
            if b > a and d > c and m.build_item_rows_fused((a, b), self.tables[0], (c, d), self.tables[1]):
                return self.tables                                 # in-vocab cast folded into the OOV embed launch
            for i, ((lo, hi), out) in enumerate(zip(self.segments, self.tables)):
                if hi <= lo:
                    continue
                if i == 0 and iv_stream is not None:
                    with torch.cuda.stream(iv_stream):
                        m.build_item_table(self.n_total, row_range=(lo, hi), out=out)
                else:
                    m.build_item_table(self.n_total, row_range=(lo, hi), out=out)
            return self.tables
        self.tables = [self._build(lo, hi) for lo, hi in self.segments]
        return self.tables

    def _local_seg(self, seg) -> Optional[Tuple[int, int]]:
        """Global id filter [a, b) as ONE range of local rows of the fused table, or None if it is not contiguous there."""
        (lo0, hi0), (lo1, hi1) = self.segments
        n0 = hi0 - lo0
        a, b = seg
        a0, b0 = max(a, lo0), min(b, hi0)
        a1, b1 = max(a, lo1), min(b, hi1)
        e0, e1 = b0 > a0, b1 > a1
        if e0 and e1:
            return (a0 - lo0, n0 + b1 - lo1) if (b0 == hi0 and a1 == lo1) else None
        if e0:
            return (a0 - lo0, b0 - lo0)
        if e1:
            return (n0 + a1 - lo1, n0 + b1 - lo1)
        return (0, 0)

    def local_history_csr(self, hist_pairs, Q: int):
        """History pairs -> CSR over the LOCAL rows of the fused table (items of other ranks dropped)."""
        from . import ops
        if hist_pairs is None or hist_pairs[0] is None:
            return None
        return ops.pairs_to_csr(hist_pairs[0], hist_pairs[1], Q, col_ranges=self.segments)

    def fused_keys(self, user_e: torch.Tensor, k: int, hist_pairs=None, seg=(0, INT64_MAX), local_csr=None) -> Optional[torch.Tensor]:
        """[Q, k] packed 8-byte candidates in global ids from ONE launch over the fused table (the final merge of the fused
        kernel writes them directly: no pack / id-remap ops); None if `seg` needs the per-segment path.
        `local_csr` = a CSR already built by `local_history_csr` (e.g. on another stream)."""
        from . import ops
        lseg = self._local_seg(seg)
        if lseg is None:
            return None
        if self.table is None:
            self.build_shard()
        (lo0, hi0), (lo1, hi1) = self.segments
        n0 = hi0 - lo0
        csr = local_csr if local_csr is not None else self.local_history_csr(hist_pairs, user_e.shape[0])
        return ops.fullsort_topk_keys(user_e, self.table, k, (n0, lo0, lo1), mask_pad=(lo0 == 0 and hi0 > 0), seg=lseg, hist=csr)

    def fused_candidates(self, user_e: torch.Tensor, k: int, hist_pairs=None, seg=(0, INT64_MAX), local_csr=None) -> Optional[torch.Tensor]:
        """[1, Q, k, 2] (score bits, global id) candidates of the fused table, the 16-byte format of the per-segment path
        (kept for callers that mix both); the exchange itself uses `fused_keys`."""
        keys = self.fused_keys(user_e, k, hist_pairs, seg, local_csr)
        if keys is None:
            return None
        from . import ops
        s, ids = ops.topk_merge_keys(keys.unsqueeze(0))
        return pack_candidates(s, ids).unsqueeze(0)

    def local_candidates(self, user_e: torch.Tensor, k: int, hist=None, seg=(0, INT64_MAX)) -> torch.Tensor:
        if self.tables is None:
            self.build_shard()
        packed = []
        for (lo, hi), tab in zip(self.segments, self.tables):
            s, i = self._local_topk(user_e, tab, k, lo, seg, hist)
            packed.append(pack_candidates(s, i))
        return torch.stack(packed, dim=0)                       # [S, Q, k, 2]

    def topk(self, user_e: torch.Tensor, k: int, hist=None, seg=(0, INT64_MAX), hist_pairs=None, local_csr=None):
        """Global (scores [Q,k], ids [Q,k]) — identical on every rank.  History either as a CSR over GLOBAL item ids
        (`hist`, per-segment path) or as the dataloader's (row, item) pairs (`hist_pairs`, fused one-launch path)."""
        if self.fused and hist is None:
            keys = self.fused_keys(user_e, k, hist_pairs, seg, local_csr)
            if keys is not None:
                from . import ops
                if self.world > 1:
                    cand = torch.empty((self.world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
                    dist.all_gather_into_tensor(cand, keys, group=self.group)      # 8 bytes per candidate
                else:
                    cand = keys.unsqueeze(0)
                return ops.topk_merge_keys(cand)
        if hist is None and hist_pairs is not None and hist_pairs[0] is not None:
            from . import ops
            hist = ops.pairs_to_csr(hist_pairs[0], hist_pairs[1], user_e.shape[0])
        local = self.local_candidates(user_e, k, hist, seg)
        if self.world > 1:
            cand = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                               device=local.device)              # ranks concatenated along dim 0
            dist.all_gather_into_tensor(cand, local, group=self.group)
        else:
            cand = local
        cs, ci = unpack_candidates(cand)
        return self._merge(cs, ci)
