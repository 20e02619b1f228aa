"""`slsh` embedder — mirrors reference inductive/single_lsh_embedder.py:9-115.

ceil(log2(n_buckets)) planes; bucket = (2 ** H).sum(1) % n_buckets with H in {0,1}, i.e.
(bits_req + popcount(bits)) % n_buckets — NOT a packed integer (single_lsh_embedder.py:82-87);
the embedding is the single row `model.*_oov_buckets(bucket)`.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .abstract_embedder import AbstractInductiveEmbedder
from .lsh_embedder import build_feature_mats
from .torch_hash import TorchLSHash


class SingleLSHInductiveEmbedder(AbstractInductiveEmbedder):
    def __init__(self, user_features, item_features, n_original_users, n_original_items, n_user_oov_buckets,
                 n_item_oov_buckets, embedding_size, device, prime_pad, normalization_type) -> None:
        super().__init__(user_features, item_features)
        self.n_original_users = n_original_users
        self.n_original_items = n_original_items
        self.n_user_oov_buckets = n_user_oov_buckets
        self.n_item_oov_buckets = n_item_oov_buckets
        self.embedding_size = embedding_size
        self.device = device
        self.prime_pad = prime_pad
        # NB: 'global' builds un-normalised matrices here, exactly like the reference (:66-75)
        self.user_feature_mat, self.item_feature_mat = build_feature_mats(self, normalization_type, device)
        self.user_bits_req = int(np.ceil(np.log2(self.n_user_oov_buckets)))
        self.item_bits_req = int(np.ceil(np.log2(self.n_item_oov_buckets)))
        self.user_lsh = TorchLSHash(hash_size=self.user_bits_req, input_dim=self.user_feature_mat.size(1), device=device)
        self.item_lsh = TorchLSHash(hash_size=self.item_bits_req, input_dim=self.item_feature_mat.size(1), device=device)
        self.tie_count = None

    def _side(self, side):
        if side == "user":
            return self.user_lsh, self.user_feature_mat, self.n_user_oov_buckets
        return self.item_lsh, self.item_feature_mat, self.n_item_oov_buckets

    def _hash_node(self, nodes, lsh, feature_mat, n_buckets) -> torch.Tensor:
        """int64 bucket ids (single_lsh_embedder.py:82-87)."""
        return ops.slsh_embed(feature_mat, lsh.uniform_planes[0].data, n_buckets, None, nodes, tie_count=self.tie_count)

    def _hash_users(self, users):
        return self._hash_node(users, self.user_lsh, self.user_feature_mat, self.n_user_oov_buckets)

    def _hash_items(self, items):
        return self._hash_node(items, self.item_lsh, self.item_feature_mat, self.n_item_oov_buckets)

    def assemble_rows(self, side, ids, model, n_old, iv_table, out=None, out_dtype=torch.float32):
        lsh, fm, nb = self._side(side)
        w = (model.user_oov_buckets if side == "user" else model.item_oov_buckets).weight.detach()
        return ops.slsh_embed(fm, lsh.uniform_planes[0].data, nb, w, ids, out=out, out_dtype=out_dtype, n_old=n_old,
                              iv_table=iv_table, prime_pad=self.prime_pad if self.training else 0,
                              tie_count=self.tie_count)

    # training: out = W[bucket]; the backward scatters g into the bucket rows the forward picked
    def train_params(self, side, model):
        return [(model.user_oov_buckets if side == "user" else model.item_oov_buckets).weight]

    def assemble_rows_train(self, side, ids, model, n_old, iv_table):
        lsh, fm, nb = self._side(side)
        w = self.train_params(side, model)[0].detach()
        out, buckets = ops.slsh_embed(fm, lsh.uniform_planes[0].data, nb, w, ids, n_old=n_old, iv_table=iv_table,
                                      prime_pad=self.prime_pad if self.training else 0, tie_count=self.tie_count, return_buckets=True)
        return out, buckets                        # bucket = -1 for in-vocab rows

    def backward_rows(self, side, saved, g, ids, n_old, model):
        w = self.train_params(side, model)[0]
        return [ops.scatter_add_rows(g, saved, torch.zeros_like(w, dtype=torch.float32))]

    def embed_user_ids(self, user_ids, model) -> torch.Tensor:
        self._depad_inplace(user_ids, self.prime_pad)
        return self.assemble_rows("user", user_ids, model, 0, None)

    def embed_item_ids(self, item_ids, model) -> torch.Tensor:
        self._depad_inplace(item_ids, self.prime_pad)
        return self.assemble_rows("item", item_ids, model, 0, None)

    def embed_all_items(self, item_embeddings, model):
        raise NotImplementedError()
