"""Retrieval models on the OOV path — mirrors reference model/abstract_recommender.py:117-203
(InductiveGeneralRecommender), model/general_recommender/bpr.py:32-163 and directau.py:18-198.

Same constructor `(config, dataset, inductive_mapper=None, inductive_embedder=None)`, parameter
names (`user_embedding`, `item_embedding`, `user_oov_buckets`, `item_oov_buckets`, hence the same
state_dict keys incl. `inductive_embedder.*`) and methods (`get_user_embedding`,
`get_item_embedding`, `ind_full_sort_predict`, `full_sort_predict`, `predict`).  Differences are in
HOW: in-vocab gather + OOV embed is one fused pass (`assemble_rows`), and `full_sort_topk` fuses
scoring, the pad/history/segment masks and top-k so the [Q, N] score matrix is never written.
Training: `calculate_loss` runs the same assemble kernels under `torch.autograd` (`_AssembleTrain`):
the backward scatters the row gradients into the in-vocab table and into the OOV buckets (`oov_scatter_add_rows`,
`oov_lsh_embed_backward`), which is what trainer.py:1748-1837 (`_train_oov`) needs from the model.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .._lib import INT64_MAX


def xavier_normal_initialization(module):
    """reference model/init.py: xavier_normal_ for Embedding / Linear weights, zero Linear bias."""
    if isinstance(module, nn.Embedding):
        nn.init.xavier_normal_(module.weight.data)
    elif isinstance(module, nn.Linear):
        nn.init.xavier_normal_(module.weight.data)
        if module.bias is not None:
            nn.init.constant_(module.bias.data, 0)


def _cfg(config, key, default=None):
    try:
        v = config[key]
    except (KeyError, IndexError):
        v = None
    return default if v is None else v


class _AssembleTrain(torch.autograd.Function):
    """rows = assemble(ids) with gradients: in-vocab rows -> d table[id] += g (nn.Embedding backward, bpr.py:56-58), OOV
    rows -> the embedder's `backward_rows` (lsh: H^T (g / |H|); slsh / mapper: scatter into the bucket rows; zero: none).
    `table` and `params` are passed as inputs so autograd routes the gradients to the Parameters."""

    @staticmethod
    def forward(ctx, model, side, ids, table, *params):
        n_old = model.n_users if side == "user" else model.n_items
        emb, mapper = model.inductive_embedder, model.inductive_mapper
        ids = ids.to(model.device).contiguous()
        if mapper is not None:
            ids = mapper.map_user_ids(ids) if side == "user" else mapper.map_item_ids(ids)
        if emb is not None:
            out, saved = emb.assemble_rows_train(side, ids, model, n_old, table.detach())
        else:
            out = torch.zeros((ids.shape[0], model.embedding_size), dtype=torch.float32, device=model.device)
            ops.gather_rows(table.detach(), ids, out=out)
            ops.gather_rows(params[0].detach(), ids, idx_offset=-n_old, out=out)
            saved = None
        ctx.model, ctx.side, ctx.n_old, ctx.ids, ctx.saved = model, side, n_old, ids, saved
        ctx.table_shape = tuple(table.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        model, side, n_old, ids = ctx.model, ctx.side, ctx.n_old, ctx.ids
        g = g.contiguous().float()
        d_table = None
        if ctx.needs_input_grad[3]:
            d_table = ops.scatter_add_rows(g, ids, torch.zeros(ctx.table_shape, dtype=torch.float32, device=g.device))
        n_params = len(ctx.needs_input_grad) - 4
        d_params = [None] * n_params
        if n_params and any(ctx.needs_input_grad[4:]):
            if model.inductive_embedder is not None:
                d_params = list(model.inductive_embedder.backward_rows(side, ctx.saved, g, ids, n_old, model))
            else:                                   # mapper only: *_oov_buckets(id - n_old), bpr.py:71
                w = (model.user_oov_buckets if side == "user" else model.item_oov_buckets).weight
                d_params = [ops.scatter_add_rows(g, ids, torch.zeros_like(w, dtype=torch.float32), idx_offset=-n_old)]
            d_params = [d if need else None for d, need in zip(d_params, ctx.needs_input_grad[4:])]
        return (None, None, None, d_table, *d_params)


class InductiveGeneralRecommender(nn.Module):
    def __init__(self, config, dataset, inductive_mapper=None, inductive_embedder=None):
        super().__init__()
        self.USER_ID = config["USER_ID_FIELD"]
        self.ITEM_ID = config["ITEM_ID_FIELD"]
        self.NEG_ITEM_ID = _cfg(config, "NEG_PREFIX", "neg_") + self.ITEM_ID
        self.n_users = dataset.num(self.USER_ID)
        self.n_items = dataset.num(self.ITEM_ID)
        self.device = config["device"]

        self.n_user_oov_buckets = 0
        self.n_item_oov_buckets = 0
        self.embedding_size = config["embedding_size"]
        self.inductive_mapper = inductive_mapper
        self.inductive_embedder = inductive_embedder
        self.oov_freeze_embedding = _cfg(config, "oov_freeze_embedding", False)
        self.oov_training = False
        if self.inductive_mapper is None and self.inductive_embedder is None:
            raise NotImplementedError("Must provide either self.inductive_mapper or self.inductive_embedder")
        self.n_new_items = self.inductive_mapper.n_new_items if self.inductive_mapper else self.inductive_embedder.n_new_items

        if _cfg(config, "add_oov_buckets", False):
            self.n_user_oov_buckets = config["user_oov_buckets"]
            self.user_oov_buckets = nn.Embedding(self.n_user_oov_buckets, self.embedding_size)
            self.n_item_oov_buckets = config["item_oov_buckets"]
            self.item_oov_buckets = nn.Embedding(self.n_item_oov_buckets, self.embedding_size)

        # dtype of the assembled tables fed to the scoring kernel ('float32' | 'bfloat16')
        td = _cfg(config, "table_dtype", "float32")
        self.table_dtype = torch.bfloat16 if str(td) in ("bfloat16", "bf16", "torch.bfloat16") else torch.float32
        self._item_cache: Optional[Tuple[int, torch.Tensor]] = None

    # --- train/eval switches (abstract_recommender.py:147-171) -----------------------------
    def set_oov_train(self, no_freeze=False):
        self.oov_training = True
        for m in (self.inductive_mapper, self.inductive_embedder):
            if m is not None:
                m.set_train()
        if self.oov_freeze_embedding and not no_freeze:
            self.freeze_non_oov_layers()

    def set_oov_eval(self, no_freeze=False):
        self.oov_training = False
        for m in (self.inductive_mapper, self.inductive_embedder):
            if m is not None:
                m.set_eval()
        if self.oov_freeze_embedding and not no_freeze:
            self.unfreeze_non_oov_layers()

    def freeze_non_oov_layers(self):
        self.user_embedding.weight.requires_grad = False
        self.item_embedding.weight.requires_grad = False

    def unfreeze_non_oov_layers(self):
        self.user_embedding.weight.requires_grad = True
        self.item_embedding.weight.requires_grad = True

    def _user_id_lookup(self, user_ids):
        return ops.gather_rows(self.user_embedding.weight.detach(), user_ids)

    def _item_id_lookup(self, item_ids):
        return ops.gather_rows(self.item_embedding.weight.detach(), item_ids)

    # --- fused gather + OOV embed (bpr.py:48-78, 94-125) ----------------------------------------
    def _assemble(self, side: str, ids: torch.Tensor, out=None, out_dtype=torch.float32) -> torch.Tensor:
        ids = ids.to(self.device)
        table = (self.user_embedding if side == "user" else self.item_embedding).weight.detach()
        n_old = self.n_users if side == "user" else self.n_items
        if self.inductive_mapper is not None:
            ids = self.inductive_mapper.map_user_ids(ids) if side == "user" else self.inductive_mapper.map_item_ids(ids)
        if self.inductive_embedder is not None:
            return self.inductive_embedder.assemble_rows(side, ids, self, n_old, table, out=out, out_dtype=out_dtype)
        # mapper only: in-vocab rows from the table, mapped OOV rows from *_oov_buckets(id - n_old)
        if out is None:
            out = torch.zeros((ids.shape[0], self.embedding_size), dtype=out_dtype, device=self.device)
        ops.gather_rows(table, ids, out=out)                                    # skips ids >= n_old
        buckets = (self.user_oov_buckets if side == "user" else self.item_oov_buckets).weight.detach()
        ops.gather_rows(buckets, ids, idx_offset=-n_old, out=out)               # skips ids < n_old
        return out

    def _assemble_train(self, side: str, ids: torch.Tensor) -> torch.Tensor:
        """fp32 rows carrying autograd history (used by `calculate_loss`)."""
        table = (self.user_embedding if side == "user" else self.item_embedding).weight
        if self.inductive_embedder is not None:
            params = self.inductive_embedder.train_params(side, self)
        else:
            params = [(self.user_oov_buckets if side == "user" else self.item_oov_buckets).weight]
        return _AssembleTrain.apply(self, side, ids, table, *params)

    def get_user_embedding(self, new_user_ids):
        if self._autograd_assemble:
            return self._assemble_train("user", new_user_ids)
        return self._assemble("user", new_user_ids)

    def get_item_embedding(self, item):
        if self._autograd_assemble:
            return self._assemble_train("item", item)
        return self._assemble("item", item)

    _autograd_assemble = False       # set for the duration of calculate_loss

    def _training_forward(self, fn):
        if not torch.is_grad_enabled():
            return fn()
        self._autograd_assemble = True
        try:
            return fn()
        finally:
            self._autograd_assemble = False

    def forward(self, user, item):
        return self.get_user_embedding(user), self.get_item_embedding(item)

    def predict(self, interaction):
        user_e, item_e = self.forward(interaction[self.USER_ID], interaction[self.ITEM_ID])
        return torch.mul(user_e, item_e).sum(dim=1)

    PREDICT_NORMALIZES = False       # `predict` = plain dot product (bpr.py:146-149); DirectAU L2-normalises both sides

    def pair_topk(self, row_idx: torch.Tensor, user_ids: torch.Tensor, item_ids: torch.Tensor, n_rows: int, k: int, segs=(None,)):
        """Sampled-candidate evaluation, fused: the (user, item) pairs of a NegSampleEvalDataLoader batch (`row_idx[p]` = the
        pair's batch user, trainer.py:547-564) -> per batch user the top-k of ITS candidates, without `predict` scores in
        memory or the [users, N] matrix of -inf.  One CSR build, one embed of the candidate items (in CSR order), one embed of
        the batch users, one scoring kernel, one selection kernel per segment in `segs` ((lo, hi) item-id ranges, None = all).
        Returns [(scores [n_rows, k], ids [n_rows, k]) per segment]; missing slots are (-inf, -1).  Every row index must
        lie in [0, n_rows) (the dataloader numbers the batch users 0 .. n_rows - 1)."""
        dev = self.device
        row_idx, user_ids, item_ids = row_idx.to(dev), user_ids.to(dev), item_ids.to(dev)
        rowptr, cols = ops.pairs_to_csr(row_idx, item_ids, n_rows, zero_tail=True)
        users_of_row = torch.zeros(n_rows, dtype=torch.int64, device=dev)
        users_of_row.index_copy_(0, row_idx, user_ids)               # all pairs of a row carry the same user id
        user_e = self._assemble("user", users_of_row, out_dtype=torch.float32)
        item_e = self._assemble("item", cols.to(torch.int64), out_dtype=torch.float32)
        out, keys = [], None
        for seg in segs:
            s, i, keys = ops.pair_topk(user_e, item_e, rowptr, cols, k, seg=seg, keys=keys, normalize=self.PREDICT_NORMALIZES)
            out.append((s, i))
        return out

    # --- full-sort: reference-shaped (dense) and fused ------------------------------------------
    def ind_full_sort_predict(self, interaction, item_ids):
        """bpr.py:151-156: flat [Q * N] raw dot products over `item_ids` (dense, for drop-in use)."""
        user_e = self._assemble("user", interaction[self.USER_ID], out_dtype=self.table_dtype)
        all_item_e = self._assemble("item", item_ids, out_dtype=self.table_dtype)
        return ops.fullsort_scores(user_e, all_item_e).view(-1)

    def full_sort_predict(self, interaction):
        user_e = self._assemble("user", interaction[self.USER_ID], out_dtype=self.table_dtype)
        all_item_e = self.item_embedding.weight.detach().to(self.table_dtype)
        return ops.fullsort_scores(user_e, all_item_e).view(-1)

    def build_item_table(self, n_total_items: Optional[int] = None, row_range: Optional[Tuple[int, int]] = None,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """All-item embedding table `get_item_embedding(arange(N))` (bpr.py:154) in `table_dtype`;
        `row_range=(lo, hi)` builds one contiguous shard (SURVEY §8e row-sharding).

        The ids are a known contiguous range, so the in-vocab / OOV split at n_items is done on the host
        (no mask, no sync): rows < n_items are a gather-cast of the table, rows >= n_items go to the
        embedder with nothing to skip."""
        n_total = self.n_new_items if n_total_items is None else n_total_items
        lo, hi = (0, n_total) if row_range is None else row_range
        n = max(hi - lo, 0)
        if out is None:
            out = torch.empty((n, self.embedding_size), dtype=self.table_dtype, device=self.device)
        if self.inductive_mapper is not None or self.inductive_embedder is None or n == 0:
            ids = torch.arange(lo, hi, device=self.device, dtype=torch.int64)
            return self._assemble("item", ids, out=out, out_dtype=self.table_dtype)
        split = min(max(self.n_items, lo), hi)
        if lo < split < hi and self.build_item_rows_fused((lo, split), out[: split - lo], (split, hi), out[split - lo:]):
            return out
        if split > lo:
            ids_iv = self._id_range(lo, split)
            ops.gather_rows(self.item_embedding.weight.detach(), ids_iv, out=out[: split - lo])
        if hi > split:
            ids_oov = self._id_range(split, hi)
            kw = {"id_range": (split, hi)} if getattr(self.inductive_embedder, "_planes_cache", None) is not None else {}
            self.inductive_embedder.assemble_rows("item", ids_oov, self, 0, None, out=out[split - lo:],
                                                  out_dtype=self.table_dtype, **kw)
        return out

    def build_item_rows_fused(self, iv_range: Tuple[int, int], iv_out: torch.Tensor, oov_range: Tuple[int, int],
                              oov_out: torch.Tensor) -> bool:
        """In-vocab rows [iv_range) and OOV rows [oov_range) of the all-item table in ONE embedder launch: the in-vocab
        part of a bf16 table is a contiguous fp32 -> bf16 cast of `item_embedding.weight` rows, which an embedder with
        `SIDE_CAST` (lsh) runs inside its own kernel instead of as a separate memory-bound launch in front of it.
        Returns False (nothing done) when the combination does not apply."""
        emb, w = self.inductive_embedder, self.item_embedding.weight.detach()
        (a, b), (c, d) = iv_range, oov_range
        if emb is None or self.inductive_mapper is not None or not getattr(emb, "SIDE_CAST", False):
            return False
        if not (0 <= a < b <= self.n_items <= c < d) or self.table_dtype != torch.bfloat16 or w.dtype != torch.float32:
            return False
        D = self.embedding_size
        if not (w.is_contiguous() and iv_out.is_contiguous() and iv_out.dtype == torch.bfloat16 and iv_out.shape == (b - a, D)):
            return False
        src = w[a:b]
        if ((b - a) * D) % 8 or src.data_ptr() % 16 or iv_out.data_ptr() % 16:
            return False
        emb.assemble_rows("item", self._id_range(c, d), self, 0, None, out=oov_out, out_dtype=self.table_dtype, side_cast=(src, iv_out))
        return True

    def _id_range(self, lo: int, hi: int) -> torch.Tensor:
        """arange(lo, hi) on the device, kept for the few ranges a model is asked for again and again (the halves of its
        item table / of its shard): saves two elementwise launches and 8 B per row of writes per step."""
        cache = self.__dict__.setdefault("_id_range_cache", {})
        key = (lo, hi, str(self.device))
        t = cache.get(key)
        if t is None:
            t = torch.arange(lo, hi, device=self.device, dtype=torch.int64)
            if len(cache) < 8:                  # never evict: a captured CUDA graph may hold the address of an entry
                cache[key] = t
        return t

    def full_sort_topk(self, interaction, k: int, n_total_items: Optional[int] = None, history_index=None,
                       seg: Tuple[int, int] = (0, INT64_MAX), item_table: Optional[torch.Tensor] = None,
                       item_id_offset: int = 0, hist_csr=None):
        """Fused replacement of ind_full_sort_predict + evaluator.py:91-94 masks + collector.py:153 top-k.

        Returns (scores fp32 [Q, k], item ids int64 [Q, k]).  The item table is rebuilt on every call
        (what bpr.py:154 does) unless `item_table` (e.g. a shard built by `build_item_table`) is given.
        `history_index` is the (row, item) pair of tensors the FullSortEvalDataLoader yields."""
        users = interaction[self.USER_ID] if not isinstance(interaction, torch.Tensor) else interaction
        user_e = self._assemble("user", users, out_dtype=self.table_dtype)
        if item_table is None:
            item_table = self.build_item_table(n_total_items)
        if hist_csr is None and history_index is not None:
            hist_csr = ops.pairs_to_csr(history_index[0].to(self.device), history_index[1].to(self.device), user_e.shape[0])
        return ops.fullsort_topk(user_e, item_table, k, item_id_offset=item_id_offset, mask_pad=True, seg=seg, hist=hist_csr)


class BPR(InductiveGeneralRecommender):
    """reference model/general_recommender/bpr.py:32-163."""

    def __init__(self, config, dataset, inductive_mapper=None, inductive_embedder=None):
        super().__init__(config, dataset, inductive_mapper, inductive_embedder)
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)
        self.apply(xavier_normal_initialization)

    def calculate_loss(self, interaction):
        """bpr.py:132-144 with BPRLoss (model/loss.py: -log(1e-10 + sigmoid(pos - neg)).mean()); the three assembles run
        through `_AssembleTrain`, the per-pair dot products and the loss are a few elementwise ops on [batch, D]."""
        def fn():
            user_e, pos_e = self.forward(interaction[self.USER_ID], interaction[self.ITEM_ID])
            neg_e = self.get_item_embedding(interaction[self.NEG_ITEM_ID])
            pos_s, neg_s = torch.mul(user_e, pos_e).sum(dim=1), torch.mul(user_e, neg_e).sum(dim=1)
            return -torch.log(1e-10 + torch.sigmoid(pos_s - neg_s)).mean()
        return self._training_forward(fn)


class DirectAU(InductiveGeneralRecommender):
    """reference model/general_recommender/directau.py:18-198.  `forward`/`predict` L2-normalise,
    `ind_full_sort_predict` does NOT (directau.py:193-198) — kept as is."""

    def __init__(self, config, dataset, inductive_mapper=None, inductive_embedder=None):
        super().__init__(config, dataset, inductive_mapper, inductive_embedder)
        self.gamma = _cfg(config, "gamma", 1.0)
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)
        self.restore_user_e = None
        self.restore_item_e = None
        self.other_parameter_name = ["restore_user_e", "restore_item_e"]
        self.apply(xavier_normal_initialization)

    PREDICT_NORMALIZES = True        # directau.py:75-78,174-181

    def forward(self, user, item):
        return F.normalize(self.get_user_embedding(user), dim=-1), F.normalize(self.get_item_embedding(item), dim=-1)

    def full_sort_predict(self, interaction):
        raise NotImplementedError()      # directau.py:183-184

    @staticmethod
    def alignment(x, y, alpha=2):
        return (x - y).norm(p=2, dim=1).pow(alpha).mean()

    @staticmethod
    def uniformity(x, t=2):
        return torch.pdist(x, p=2).pow(2).mul(-t).exp().mean().log()

    def calculate_loss(self, interaction):
        """directau.py:87-99: alignment + gamma * mean uniformity of the L2-normalised rows; the assembles run through
        `_AssembleTrain`."""
        self.restore_user_e, self.restore_item_e = None, None

        def fn():
            user_e, item_e = self.forward(interaction[self.USER_ID], interaction[self.ITEM_ID])
            return self.alignment(user_e, item_e) + self.gamma * (self.uniformity(user_e) + self.uniformity(item_e)) / 2
        return self._training_forward(fn)
