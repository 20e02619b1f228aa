"""Full-sort evaluation on the fused path — mirrors reference inductive/evaluator.py:15-180
(InductiveEvaluator.eval_batch / evaluate_model), evaluator/collector.py:137-194
(Collector.eval_batch_collect, 'rec.topk') and inductive/filtered_collector.py:18-80 +
collector_filter.py:128-256 (the six old/new user x old/new item collectors).

The reference materialises scores [Q, N], masks them in place and runs torch.topk up to 7 times per batch.  Here ONE
scoring pass over the items — split at n_old_items into two fused launches, so every item is scored once — yields the
old-items-only and new-items-only top-k lists, their merge is the all-items list, and one kernel turns the three lists
and the CSR positives into the [hits | pos_len | keep] rows of all seven collectors.  Nothing in `eval_batch`
synchronises with the host (no boolean-mask indexing, no unique, no per-collector .cpu()): results stay on the device
until they are read, and `evaluate_model` copies them to the host once.  Each collector sees un-aliased scores (the
reference's in-place `-inf` writes leak from one filtered collector into the next, collector_filter.py:172-175;
SURVEY §8f row 1).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from ._lib import INT64_MAX

COLLECTORS = {          # name -> (return_old_users, return_old_items); None = no filter (evaluator.py:29-47)
    "overall": (None, None),
    "old_users": (True, None), "new_users": (False, None),
    "old_old": (True, True), "old_new": (True, False),
    "new_old": (False, True), "new_new": (False, False),
}


class Collector:
    """Accumulates 'rec.topk' rows; `eval_batch_collect` keeps the reference signature for callers
    that still hold a dense score matrix (collector.py:137-167).  Rows stay on the device until
    `get_data_struct()` is called (one host copy for all batches)."""

    def __init__(self, config):
        self.config = config
        self.topk = list(config["topk"])
        self.k = max(self.topk)
        self._rows: List[torch.Tensor] = []          # host rows already materialised
        self._pending: List = []                     # callables returning device/host rows, resolved on read

    def collect_topk(self, topk_idx: torch.Tensor, positive_u: torch.Tensor, positive_i: torch.Tensor) -> torch.Tensor:
        Q = topk_idx.shape[0]
        rowptr, cols = ops.pairs_to_csr(positive_u.to(topk_idx.device), positive_i.to(topk_idx.device), Q)
        res = ops.topk_hits(topk_idx, rowptr, cols)
        self._pending.append(lambda r=res: r)
        return res

    def eval_batch_collect(self, scores_tensor: torch.Tensor, interaction, positive_u, positive_i):
        """Dense-score entry point: top-k over an already-masked [Q, N] matrix."""
        _, topk_idx = ops.dense_topk(scores_tensor, self.k)
        return self.collect_topk(topk_idx, positive_u, positive_i)

    def get_data_struct(self) -> Dict[str, torch.Tensor]:
        for fn in self._pending:
            r = fn()
            if r is not None and r.shape[0]:
                self._rows.append(r.cpu())
        self._pending = []
        return {"rec.topk": torch.cat(self._rows, dim=0) if self._rows else torch.zeros((0, self.k + 1), dtype=torch.int32)}


class BatchResult:
    """`{collector name: 'rec.topk' rows of this batch}` — read lazily: the first access copies the batch's
    [7, Q, k + 2] result tensor to the host (the only synchronisation) and splits it by the `keep` column.  A collector
    without rows is absent, like a reference FilteredCollector that returns early (filtered_collector.py:34-35)."""

    def __init__(self, dev_rows: torch.Tensor):
        self.device_rows = dev_rows                  # int32 [7, Q, k + 2]: hits | pos_len | keep
        self._host: Optional[Dict[str, torch.Tensor]] = None

    def _materialise(self) -> Dict[str, torch.Tensor]:
        if self._host is None:
            rows = self.device_rows.cpu()
            out = {}
            for c, name in enumerate(COLLECTORS):
                keep = rows[c, :, -1] != 0
                if name == "overall" or bool(keep.any()):
                    out[name] = rows[c][keep][:, :-1].contiguous()
            self._host = out
        return self._host

    def rows_of(self, name: str) -> Optional[torch.Tensor]:
        return self._materialise().get(name)

    def __contains__(self, name): return name in self._materialise()
    def __getitem__(self, name): return self._materialise()[name]
    def __iter__(self): return iter(self._materialise())
    def keys(self): return self._materialise().keys()
    def items(self): return self._materialise().items()
    def __len__(self): return len(self._materialise())


def topk_metrics(rec_topk: np.ndarray, topk: Sequence[int]) -> Dict[str, float]:
    """hit / recall / precision / ndcg / mrr @k from [hits | pos_len] rows — the numpy maths of
    reference evaluator/metrics.py + base_metric.py:45-100 (consumer of the hot path; tiny CPU work)."""
    rec = np.asarray(rec_topk)
    if rec.shape[0] == 0:
        return {}
    pos_idx, pos_len = rec[:, :-1].astype(bool), rec[:, -1].astype(np.int64)
    out = {}
    kmax = pos_idx.shape[1]
    cum = np.cumsum(pos_idx, axis=1)
    ranks = np.arange(1, kmax + 1)
    dcg = np.cumsum(np.where(pos_idx, 1.0 / np.log2(ranks + 1), 0.0), axis=1)
    ideal = np.cumsum(1.0 / np.log2(ranks + 1))
    first = np.where(pos_idx.any(1), pos_idx.argmax(1) + 1, 0)
    for k in topk:
        k = min(k, kmax)
        hits_k = cum[:, k - 1]
        out[f"hit@{k}"] = float((hits_k > 0).mean())
        out[f"recall@{k}"] = float((hits_k / np.maximum(pos_len, 1)).mean())
        out[f"precision@{k}"] = float((hits_k / k).mean())
        idcg = ideal[np.minimum(np.maximum(pos_len, 1), k) - 1]
        out[f"ndcg@{k}"] = float((dcg[:, k - 1] / idcg).mean())
        out[f"mrr@{k}"] = float(np.where((first > 0) & (first <= k), 1.0 / np.maximum(first, 1), 0.0).mean())
    return out


class InductiveEvaluator:
    """Drop-in for reference inductive/evaluator.py: same constructor, `eval_batch`, `evaluate_model`.

    reference_compat=False (default): a collector with an item filter reads the k-list of ITS item segment
    (return_old_items) and hits are compared in global item ids.
    reference_compat=True reproduces two quirks of collector_filter.py bit for bit (opt-in; the golden fixtures of the
    reference's FilteredCollector are checked in this mode):
      * :172-175 picks the masked item segment from `return_old_USERS` (old users -> new items blanked, new users ->
        old items blanked), whatever `return_old_items` says;
      * :249-250 shifts new-item positives by -n_old_items while the score columns stay global.
    It does not reproduce the in-place -inf leak from one collector into the next (each collector sees clean scores)."""

    def __init__(self, model, config, n_old_users, n_old_items, feature_extractor=None, reference_compat=False):
        self.model = model
        self.config = config
        self.device = model.device
        self.USER_ID = config["USER_ID_FIELD"]
        self.ITEM_ID = config["ITEM_ID_FIELD"]
        self.n_old_users, self.n_old_items = n_old_users, n_old_items
        self.topk = list(config["topk"])
        self.k = max(self.topk)
        self.collectors = {name: Collector(config) for name in COLLECTORS}
        self.tot_item_num: Optional[int] = None
        self.item_range = None
        self.reference_compat = reference_compat

    def eval_batch(self, batched_data, item_table: Optional[torch.Tensor] = None) -> BatchResult:
        """(interaction, history_index, positive_u, positive_i) -> {collector: 'rec.topk' rows of this batch} (lazy).
        Launches: two CSR builds, the user embed, two fused score + mask + top-k launches that together visit every item
        once (old items, new items), one merge, one 7-collector hits kernel.  No host synchronisation."""
        interaction, history_index, positive_u, positive_i = batched_data
        users = interaction[self.USER_ID].to(self.device)
        Q = users.shape[0]
        hist = None
        if history_index is not None:
            hist = ops.pairs_to_csr(history_index[0].to(self.device), history_index[1].to(self.device), Q)
        if item_table is None:
            item_table = self.model.build_item_table(self.tot_item_num)
        m = self.model
        user_e = m._assemble("user", users, out_dtype=m.table_dtype)
        s_old, i_old = ops.fullsort_topk(user_e, item_table, self.k, mask_pad=True, seg=(0, self.n_old_items), hist=hist)
        s_new, i_new = ops.fullsort_topk(user_e, item_table, self.k, mask_pad=True, seg=(self.n_old_items, INT64_MAX), hist=hist)
        # the two segments partition the items, so the all-items list is the merge of their lists (same total order)
        _, i_all = ops.topk_merge(torch.stack([s_old, s_new]), torch.stack([i_old, i_new]))
        rowptr, cols = ops.pairs_to_csr(positive_u.to(self.device), positive_i.to(self.device), Q)
        rows = ops.topk_hits_collectors(i_all, i_old, i_new, users, self.n_old_users, self.n_old_items, rowptr, cols,
                                        reference_compat=self.reference_compat)
        res = BatchResult(rows)
        for name in COLLECTORS:
            self.collectors[name]._pending.append(lambda r=res, n=name: r.rows_of(n))
        return res

    def neg_sample_batch_eval(self, batched_data) -> BatchResult:
        """Sampled-negative evaluation (`eval_args.mode: uni250` etc.; reference inductive/evaluator.py:116-133 +
        the seven collectors): (interaction with USER_ID / ITEM_ID pairs, row_idx, positive_u, positive_i) ->
        {collector: 'rec.topk' rows} (lazy).  The reference scatters `model.predict` into a [users, N] matrix of -inf and
        runs topk per collector; here every pair is scored once, the old-items and new-items lists come from the same keys
        and the all-items list is their merge.  The number of batch users must be passed by the dataloader as
        `positive_u[-1] + 1` on the HOST (like trainer.py:559 reads it) or as interaction["n_rows"]."""
        interaction, row_idx, positive_u, positive_i = batched_data
        users, items = interaction[self.USER_ID], interaction[self.ITEM_ID]
        try:
            n_rows = int(interaction["n_rows"])
        except (KeyError, IndexError, TypeError):
            n_rows = int(positive_u[-1]) + 1                         # trainer.py:559 (a host read, as in the reference)
        (s_old, i_old), (s_new, i_new) = self.model.pair_topk(row_idx, users, items, n_rows, self.k,
                                                              segs=((0, self.n_old_items), (self.n_old_items, INT64_MAX)))
        _, i_all = ops.topk_merge(torch.stack([s_old, s_new]), torch.stack([i_old, i_new]))
        dev = self.device
        users_of_row = torch.zeros(n_rows, dtype=torch.int64, device=dev)
        users_of_row.index_copy_(0, row_idx.to(dev), users.to(dev))
        rowptr, cols = ops.pairs_to_csr(positive_u.to(dev), positive_i.to(dev), n_rows)
        rows = ops.topk_hits_collectors(i_all, i_old, i_new, users_of_row, self.n_old_users, self.n_old_items, rowptr, cols,
                                        reference_compat=self.reference_compat)
        res = BatchResult(rows)
        for name in COLLECTORS:
            self.collectors[name]._pending.append(lambda r=res, n=name: r.rows_of(n))
        return res

    def evaluate_model(self, eval_data, config=None, show_progress=False, inductive=True, n_total_items=None):
        self.model.eval()
        self.tot_item_num = n_total_items if n_total_items is not None else eval_data._dataset.item_num
        # weights are frozen during evaluation: build the item table once instead of once per batch (bpr.py:154)
        item_table = self.model.build_item_table(self.tot_item_num)
        for batched_data in eval_data:
            self.eval_batch(batched_data, item_table=item_table)
        return {name: topk_metrics(c.get_data_struct()["rec.topk"].numpy(), self.topk) for name, c in self.collectors.items()}
