"""Token gather + OOV overwrite of the ranking models — mirrors reference
model/abstract_recommender.py:715-842 (InductiveContextRecommender.embed_token_fields),
model/layers.py:130-153 (FMEmbedding) and :1617-1750 (InductiveFMFirstOrderLinear).

Only the gather/OOV-overwrite step is on the accelerated path; the dense towers of
DCNV2 / WideDeep / xDeepFM that consume these tensors are "next" (§8f row 2) and stay the caller's.
Column 0 of `token_fields` is the user id, column 1 the item id (abstract_recommender.py:691-692).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from torch import nn

from .. import ops


class FMEmbedding(nn.Module):
    """layers.py:130-153: one table for all token fields, per-field row offsets."""

    def __init__(self, field_dims: Sequence[int], offsets, embed_dim: int):
        super().__init__()
        self.embedding = nn.Embedding(int(sum(field_dims)), embed_dim)
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self._offsets_dev: Optional[torch.Tensor] = None

    def offsets_tensor(self, device) -> torch.Tensor:
        if self._offsets_dev is None or self._offsets_dev.device != torch.device(device):
            self._offsets_dev = torch.as_tensor(self.offsets, dtype=torch.int64, device=device)
        return self._offsets_dev

    def forward(self, input_x: torch.Tensor) -> torch.Tensor:
        w = self.embedding.weight.detach()
        return ops.token_gather(input_x, self.offsets_tensor(w.device), w, n_users=1 << 62, n_items=1 << 62)


class _TokenOOVMixin:
    """Shared OOV-overwrite logic of abstract_recommender.py:794-842 and layers.py:1634-1693."""

    def _embed_tokens(self, token_fields: torch.Tensor, uid_idx: int, iid_idx: int) -> torch.Tensor:
        w = self.token_embedding_table.embedding.weight.detach()
        dev = w.device
        token_fields = token_fields.to(dev)
        offsets = self.token_embedding_table.offsets_tensor(dev)
        fields = token_fields.shape[1]
        D = w.shape[1]
        emb, mapper = self.inductive_embedder, self.inductive_mapper
        if mapper is None and emb is None:
            raise RuntimeError("Must provide either self.inductive_mapper or self.inductive_embedder")
        # 1) every in-vocab cell; OOV user/item cells are left for step 2 (only the overwrite is observable)
        out = ops.token_gather(token_fields, offsets, w, self.n_users, self.n_items, uid_idx=uid_idx, iid_idx=iid_idx)
        # 2) OOV cells, written in place through a strided view (ids_stride = fields, out_stride = fields * D)
        for side, col, n_old in (("user", uid_idx, self.n_users), ("item", iid_idx, self.n_items)):
            ids = token_fields[:, col]
            view = out[:, col, :]
            if mapper is not None:
                mapped = mapper.map_user_ids(ids.contiguous()) if side == "user" else mapper.map_item_ids(ids.contiguous())
                buckets = (self.user_oov_buckets if side == "user" else self.item_oov_buckets).weight.detach()
                ops.gather_rows(buckets, mapped, idx_offset=-n_old, out=view)   # in-vocab ids (< n_old) are skipped
            else:
                emb.assemble_rows(side, ids, self, n_old, None, out=view, out_dtype=out.dtype)
        return out


class InductiveContextRecommender(nn.Module, _TokenOOVMixin):
    """The OOV-aware embedding front-end of DCNV2 / WideDeep / xDeepFM.

    `field_dims[0]` / `[1]` are the user / item vocabularies (n_users, n_items)."""

    def __init__(self, config, field_dims: Sequence[int], inductive_mapper=None, inductive_embedder=None,
                 first_order_embedder=None, first_order_mapper=None):
        super().__init__()
        self.embedding_size = config["embedding_size"]
        self.inductive_mapper = inductive_mapper
        self.inductive_embedder = inductive_embedder
        self.oov_training = False
        self.n_users, self.n_items = int(field_dims[0]), int(field_dims[1])
        if inductive_mapper is None and inductive_embedder is None:
            raise NotImplementedError("Must provide either self.inductive_mapper or self.inductive_embedder")
        offsets = np.array((0, *np.cumsum(field_dims)[:-1]), dtype=np.int64)
        self.token_field_offsets = offsets
        self.token_embedding_table = FMEmbedding(field_dims, offsets, self.embedding_size)
        try:
            add = config["add_oov_buckets"]
        except (KeyError, IndexError):
            add = False
        if add:
            self.n_user_oov_buckets = config["user_oov_buckets"]
            self.user_oov_buckets = nn.Embedding(self.n_user_oov_buckets, self.embedding_size)
            self.n_item_oov_buckets = config["item_oov_buckets"]
            self.item_oov_buckets = nn.Embedding(self.n_item_oov_buckets, self.embedding_size)
        if first_order_embedder is not None or first_order_mapper is not None:
            # abstract_recommender.py:748-760: a second, independently-initialised embedder with embedding_size = 1
            self.first_order_linear = InductiveFMFirstOrderLinear(config, field_dims, self.n_users, self.n_items,
                                                                  inductive_mapper=first_order_mapper,
                                                                  inductive_embedder=first_order_embedder)

    def embed_token_fields(self, token_fields: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """[B, fields] int64 -> [B, fields, D] (abstract_recommender.py:794-842)."""
        if token_fields is None:
            return None
        return self._embed_tokens(token_fields, 0, 1)


class InductiveFMFirstOrderLinear(nn.Module, _TokenOOVMixin):
    """layers.py:1617-1750 restricted to the token fields: D = 1 table + D = 1 OOV buckets."""

    def __init__(self, config, field_dims: Sequence[int], n_users: int, n_items: int, output_dim: int = 1,
                 inductive_mapper=None, inductive_embedder=None):
        super().__init__()
        self.n_users, self.n_items = n_users, n_items
        self.inductive_mapper = inductive_mapper
        self.inductive_embedder = inductive_embedder
        offsets = np.array((0, *np.cumsum(field_dims)[:-1]), dtype=np.int64)
        self.token_field_offsets = offsets
        self.token_embedding_table = FMEmbedding(field_dims, offsets, output_dim)
        self.bias = nn.Parameter(torch.zeros((output_dim,)), requires_grad=True)
        try:
            add = config["add_oov_buckets"]
        except (KeyError, IndexError):
            add = False
        if add:
            self.n_user_oov_buckets = config["user_oov_buckets"]
            self.user_oov_buckets = nn.Embedding(self.n_user_oov_buckets, output_dim)
            self.n_item_oov_buckets = config["item_oov_buckets"]
            self.item_oov_buckets = nn.Embedding(self.n_item_oov_buckets, output_dim)

    def embed_token_fields(self, token_fields, uid_idx=None, iid_idx=None):
        """[B, fields] -> [B, 1, output_dim]: per-field first-order weights summed over fields."""
        if token_fields is None:
            return None
        w = self.token_embedding_table.embedding.weight.detach()
        dev = w.device
        token_fields = token_fields.to(dev)
        if w.shape[1] != 1 or uid_idx is None or iid_idx is None:
            if uid_idx is None or iid_idx is None:
                e = self.token_embedding_table(token_fields)
            else:
                e = self._embed_tokens(token_fields, uid_idx, iid_idx)
            return torch.sum(e, dim=1, keepdim=True)
        # D = 1 fast path: OOV cell values into two [B] scratch vectors, then one fused sum over fields
        offsets = self.token_embedding_table.offsets_tensor(dev)
        Bn = token_fields.shape[0]
        vals = {}
        for side, col, n_old in (("user", uid_idx, self.n_users), ("item", iid_idx, self.n_items)):
            ids = token_fields[:, col]
            scratch = torch.zeros((Bn, 1), dtype=torch.float32, device=dev)
            if self.inductive_mapper is not None:
                mp = self.inductive_mapper
                mapped = mp.map_user_ids(ids.contiguous()) if side == "user" else mp.map_item_ids(ids.contiguous())
                buckets = (self.user_oov_buckets if side == "user" else self.item_oov_buckets).weight.detach()
                ops.gather_rows(buckets, mapped, idx_offset=-n_old, out=scratch)
            elif self.inductive_embedder is not None:
                self.inductive_embedder.assemble_rows(side, ids, self, n_old, None, out=scratch, out_dtype=torch.float32)
            else:
                raise RuntimeError("Must provide either self.inductive_mapper or self.inductive_embedder")
            vals[side] = scratch.view(-1)
        s = ops.first_order_sum(token_fields, offsets, w, self.n_users, self.n_items, vals["user"], vals["item"],
                                uid_idx=uid_idx, iid_idx=iid_idx)
        return s.view(Bn, 1, 1)

    def forward(self, token_fields: torch.Tensor) -> torch.Tensor:
        """Token part of layers.py:1695-1750: sum over fields + bias -> [B, output_dim]."""
        return self.embed_token_fields(token_fields, 0, 1).sum(dim=1) + self.bias
