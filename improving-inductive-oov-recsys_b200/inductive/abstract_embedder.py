"""Plugin base class — mirrors reference inductive/abstract_embedder.py:5-70.

Same constructor, attributes (`n_new_users`, `n_new_items`, `training`) and methods
(`set_train`, `set_eval`, `embed_user_ids`, `embed_item_ids`).  One addition: the fused
"assemble" entry points, which gather in-vocab rows and embed OOV rows in ONE kernel pass
(no boolean-mask indexing, no device->host sync) — what bpr.py:48-125 does in five ops.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn


def feature_columns(features):
    """Column names of a feature container: RecBole Interaction / pandas frame (`.columns`) or dict."""
    cols = getattr(features, "columns", None)
    if cols is None:
        cols = list(features.keys())
    return list(cols)


def feature_block(features, col, n: int) -> torch.Tensor:
    """`features[col].float().view(n, -1)` (lsh_embedder.py:83-90)."""
    v = features[col]
    if not isinstance(v, torch.Tensor):
        v = torch.as_tensor(v)
    return v.float().reshape(n, -1)


class AbstractInductiveEmbedder(nn.Module):
    def __init__(self, user_features, item_features) -> None:
        super().__init__()
        self.user_features = user_features
        self.item_features = item_features
        self.n_new_users = len(user_features)
        self.n_new_items = len(item_features)
        self.training = False

    def set_train(self):
        self.training = True

    def set_eval(self):
        self.training = False

    # --- reference API -------------------------------------------------------------------
    def embed_user_ids(self, user_ids: torch.LongTensor, model) -> torch.Tensor:
        raise NotImplementedError()

    def embed_item_ids(self, item_ids: torch.LongTensor, model) -> torch.Tensor:
        raise NotImplementedError()

    def map_all_item_embeddings(self, item_embeddings) -> torch.Tensor:
        raise NotImplementedError()

    # --- fused extension -----------------------------------------------------------------
    def assemble_rows(self, side: str, ids: torch.Tensor, model, n_old: int,
                      iv_table: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                      out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """out[i] = iv_table[ids[i]] if ids[i] < n_old else embed(ids[i]); `iv_table=None`
        leaves in-vocab rows of `out` untouched.  `side` is 'user' or 'item'."""
        raise NotImplementedError()

    # --- training-mode OOV path (trainer.py:1748-1837 differentiates the assemble) -----------------
    def train_params(self, side: str, model) -> list:
        """Parameters the OOV rows of `side` depend on (the autograd inputs next to the in-vocab table)."""
        return []

    def assemble_rows_train(self, side: str, ids: torch.Tensor, model, n_old: int, iv_table: torch.Tensor):
        """fp32 assemble like `assemble_rows`, plus whatever `backward_rows` needs: returns (out, saved)."""
        raise NotImplementedError(f"{type(self).__name__} has no training-mode (backward) path")

    def backward_rows(self, side: str, saved, g: torch.Tensor, ids: torch.Tensor, n_old: int, model) -> list:
        """Gradients of `train_params(side, model)` given g = d loss / d assembled rows (fp32 [n, D])."""
        return []

    def _depad_inplace(self, ids: torch.Tensor, prime_pad: int) -> None:
        """Training mode mutates the caller's ids like lsh_embedder.py:153-155 does."""
        if self.training and prime_pad:
            mask = ids >= prime_pad
            ids[mask] = ids[mask] - prime_pad
