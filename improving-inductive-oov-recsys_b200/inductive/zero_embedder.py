"""`zero` embedder — mirrors reference inductive/zero_embedder.py:6-60: OOV rows are zeros."""
from __future__ import annotations

import torch

from .. import ops
from .abstract_embedder import AbstractInductiveEmbedder


class ZeroEmbedder(AbstractInductiveEmbedder):
    def __init__(self, user_features, item_features, n_original_users, n_original_items, embedding_size, device) -> None:
        super().__init__(user_features, item_features)
        self.zero_vec = torch.zeros(embedding_size, device=device)
        self.n_original_users = n_original_users
        self.n_original_items = n_original_items

    def assemble_rows(self, side, ids, model, n_old, iv_table, out=None, out_dtype=torch.float32):
        return ops.const_embed(None, ids, self.zero_vec.numel(), out=out, out_dtype=out_dtype, n_old=n_old, iv_table=iv_table)

    def assemble_rows_train(self, side, ids, model, n_old, iv_table):      # OOV rows are constants: nothing to save
        return self.assemble_rows(side, ids, model, n_old, iv_table), None

    def embed_user_ids(self, user_ids, model) -> torch.Tensor:
        return self.assemble_rows("user", user_ids, model, 0, None)

    def embed_item_ids(self, item_ids, model) -> torch.Tensor:
        return self.assemble_rows("item", item_ids, model, 0, None)
