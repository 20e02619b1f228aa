"""`mean` embedder — mirrors reference inductive/mean_embedder.py:12-87.

OOV rows get the column mean over ALL rows (incl. pad row 0) of the in-vocab table, cached on
first use: `model.{user,item}_embedding.weight` for BPR/DirectAU, the user / item slice of the
token table for the context models (mean_embedder.py:55-60, 72-85).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from .abstract_embedder import AbstractInductiveEmbedder


class MeanEmbedder(AbstractInductiveEmbedder):
    def __init__(self, user_features, item_features, n_original_users, n_original_items, n_user_oov_buckets,
                 n_item_oov_buckets, embedding_size, device) -> None:
        super().__init__(user_features, item_features)
        self.user_feat_mean: Optional[torch.Tensor] = None
        self.item_feat_mean: Optional[torch.Tensor] = None
        self.n_original_users = n_original_users
        self.n_original_items = n_original_items

    @staticmethod
    def _table(side: str, model) -> torch.Tensor:
        if hasattr(model, "user_embedding") and hasattr(model, "item_embedding"):      # BPR / DirectAU
            return (model.user_embedding if side == "user" else model.item_embedding).weight.detach()
        if hasattr(model, "token_embedding_table"):                                     # DCNV2 / WideDeep / xDeepFM / first-order
            w = model.token_embedding_table.embedding.weight.detach()
            off = [int(x) for x in model.token_field_offsets]
            if side == "user":
                return w[off[0]:off[1]]
            return w[off[1]:] if len(off) == 2 else w[off[1]:off[2]]
        raise ValueError("Invalid model type for mean embedder")

    def _mean(self, side: str, model) -> torch.Tensor:
        cur = self.user_feat_mean if side == "user" else self.item_feat_mean
        if cur is None:
            cur = ops.col_mean(self._table(side, model))
            if side == "user":
                self.user_feat_mean = cur
            else:
                self.item_feat_mean = cur
        return cur

    def assemble_rows(self, side, ids, model, n_old, iv_table, out=None, out_dtype=torch.float32):
        vec = self._mean(side, model)
        return ops.const_embed(vec, ids, vec.numel(), out=out, out_dtype=out_dtype, n_old=n_old, iv_table=iv_table)

    def assemble_rows_train(self, side, ids, model, n_old, iv_table):
        """Training mode: the reference embeds under `torch.no_grad()` (mean_embedder.py:40, 63) — the OOV rows are constants
        (the mean cached at first use), only the in-vocab rows carry gradient."""
        return self.assemble_rows(side, ids, model, n_old, iv_table), None

    @torch.no_grad()
    def embed_user_ids(self, user_ids, model) -> torch.Tensor:
        return self.assemble_rows("user", user_ids, model, 0, None)

    @torch.no_grad()
    def embed_item_ids(self, item_ids, model) -> torch.Tensor:
        return self.assemble_rows("item", item_ids, model, 0, None)
