"""Full-sort evaluation on the fused path — mirrors reference inductive/evaluator.py:15-180
(InductiveEvaluator.eval_batch / evaluate_model), evaluator/collector.py:137-194
(Collector.eval_batch_collect, 'rec.topk') and inductive/filtered_collector.py:18-80 +
collector_filter.py:128-256 (the six old/new user x old/new item collectors).

The reference materialises scores [Q, N], masks them in place and runs torch.topk up to 7 times per
batch.  Here one fused kernel pass per item segment (all / old / new) yields top-k ids directly and
the hit matrix [hits | pos_len] is built from CSR positives — three passes serve all seven
collectors, and each collector sees un-aliased scores (the reference's in-place `-inf` writes leak
from one filtered collector into the next, collector_filter.py:172-175).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

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
    that still hold a dense score matrix (collector.py:137-167)."""

    def __init__(self, config):
        self.config = config
        self.topk = list(config["topk"])
        self.k = max(self.topk)
        self._rows = []

    def collect_topk(self, topk_idx: torch.Tensor, positive_u: torch.Tensor, positive_i: torch.Tensor) -> torch.Tensor:
        Q = topk_idx.shape[0]
        rowptr, cols = ops.pairs_to_csr(positive_u.to(topk_idx.device), positive_i.to(topk_idx.device), Q)
        res = ops.topk_hits(topk_idx, rowptr, cols)
        self._rows.append(res.cpu())                         # collector.py:44-52: results leave the GPU per batch
        return res

    def eval_batch_collect(self, scores_tensor: torch.Tensor, interaction, positive_u, positive_i):
        """Dense-score entry point: top-k over an already-masked [Q, N] matrix."""
        _, topk_idx = ops.dense_topk(scores_tensor, self.k)
        return self.collect_topk(topk_idx, positive_u, positive_i)

    def get_data_struct(self) -> Dict[str, torch.Tensor]:
        return {"rec.topk": torch.cat(self._rows, dim=0) if self._rows else torch.zeros((0, self.k + 1), dtype=torch.int32)}


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
    """Drop-in for reference inductive/evaluator.py: same constructor, `eval_batch`, `evaluate_model`."""

    def __init__(self, model, config, n_old_users, n_old_items, feature_extractor=None, reference_compat=True):
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
        # reference_compat=True reproduces two quirks of collector_filter.py bit for bit:
        #   * :172-175 picks the masked item segment from `return_old_USERS` (old users -> new items blanked,
        #     new users -> old items blanked), whatever `return_old_items` says;
        #   * :249-250 shifts new-item positives by -n_old_items while the score columns stay global.
        # reference_compat=False masks by `return_old_items` and compares in global ids.
        self.reference_compat = reference_compat

    def eval_batch(self, batched_data, item_table: Optional[torch.Tensor] = None):
        """(interaction, history_index, positive_u, positive_i) -> {collector: 'rec.topk' rows of this batch}."""
        interaction, history_index, positive_u, positive_i = batched_data
        users = interaction[self.USER_ID].to(self.device)
        Q = users.shape[0]
        hist = None
        if history_index is not None:
            hist = ops.pairs_to_csr(history_index[0].to(self.device), history_index[1].to(self.device), Q)
        if item_table is None:
            item_table = self.model.build_item_table(self.tot_item_num)
        positive_u, positive_i = positive_u.to(self.device), positive_i.to(self.device)
        passes = {}
        for seg_name, seg in (("all", (0, INT64_MAX)), ("old", (0, self.n_old_items)), ("new", (self.n_old_items, INT64_MAX))):
            passes[seg_name] = self.model.full_sort_topk(users, self.k, item_table=item_table, hist_csr=hist, seg=seg)[1]
        old_user_rows = users < self.n_old_users
        results = {}
        for name, (ru, ri) in COLLECTORS.items():
            keep_old_segment = ru if self.reference_compat else ri
            idx = passes["all" if ri is None else ("old" if keep_old_segment else "new")]
            pmask = torch.ones_like(positive_u, dtype=torch.bool)
            if ri is not None:
                pmask &= (positive_i < self.n_old_items) if ri else (positive_i >= self.n_old_items)
            if ru is not None:
                pmask &= old_user_rows[positive_u] if ru else ~old_user_rows[positive_u]
            pu, pi = positive_u[pmask], positive_i[pmask]
            if name != "overall":
                if pu.numel() == 0:
                    continue                                   # filtered_collector.py:34-35
                keep = torch.unique(pu, sorted=True)           # rows = users that still own a positive
                remap = torch.full((Q,), -1, dtype=torch.int64, device=self.device)
                remap[keep] = torch.arange(keep.numel(), device=self.device)
                pu = remap[pu]
                idx = idx[keep]
                if ri is False and self.reference_compat:
                    pi = pi - self.n_old_items
            results[name] = self.collectors[name].collect_topk(idx, pu, pi)
        return results

    def evaluate_model(self, eval_data, config=None, show_progress=False, inductive=True, n_total_items=None):
        self.model.eval()
        self.tot_item_num = n_total_items if n_total_items is not None else eval_data._dataset.item_num
        # weights are frozen during evaluation: build the item table once instead of once per batch (bpr.py:154)
        item_table = self.model.build_item_table(self.tot_item_num)
        for batched_data in eval_data:
            self.eval_batch(batched_data, item_table=item_table)
        return {name: topk_metrics(c.get_data_struct()["rec.topk"].numpy(), self.topk) for name, c in self.collectors.items()}
