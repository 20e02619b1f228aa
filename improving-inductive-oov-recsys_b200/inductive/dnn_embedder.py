"""`dnn` embedder — mirrors reference inductive/dnn_embedder.py:8-112.

A 4-layer net (Linear-GELU x3, Linear-Sigmoid, width `dhe_layer_size`) on the per-column normalised feature row of
the id; same constructor and state_dict keys (`user_hash_net.*`, `item_hash_net.*`) as the reference.  Runs through
`oov_fdhe_embed` with no hash inputs (net H = 0): feature fetch, four layers and the in-vocab / OOV assemble in one call.
Training mode de-pads the feature lookup only (dnn_embedder.py:93-109); the caller's ids are not modified.
"""
from __future__ import annotations

import torch

from .. import ops
from .abstract_embedder import AbstractInductiveEmbedder
from .dh_embedder import HashNetTraining
from .feat_dh_embedder import _feature_mats, _hash_net


class DNNEmbedder(HashNetTraining, AbstractInductiveEmbedder):
    def __init__(self, user_features, item_features, n_original_users, n_original_items, n_user_oov_buckets,
                 n_item_oov_buckets, embedding_size, device, prime_pad, dhe_layer_size) -> None:
        super().__init__(user_features, item_features)
        self.n_original_users = n_original_users
        self.n_original_items = n_original_items
        self.n_user_oov_buckets = n_user_oov_buckets
        self.n_item_oov_buckets = n_item_oov_buckets
        self.embedding_size = embedding_size
        self.device = device
        self.prime_pad = prime_pad
        self.user_feature_mat, self.item_feature_mat = _feature_mats(self, device)
        self.user_hash_net = _hash_net(self.user_feature_mat.size(1), dhe_layer_size, embedding_size, device)
        self.item_hash_net = _hash_net(self.item_feature_mat.size(1), dhe_layer_size, embedding_size, device)
        self.compute_path = ops.PATH_AUTO

    def _side(self, side: str):
        if side == "user":
            return self.user_hash_net, self.user_feature_mat
        return self.item_hash_net, self.item_feature_mat

    def assemble_rows(self, side, ids, model, n_old, iv_table, out=None, out_dtype=torch.float32):
        net, fm = self._side(side)
        return ops.fdhe_embed(ids, None, ops.DheNet.from_sequential(net, n_feat=fm.shape[1]), fm, out=out, out_dtype=out_dtype,
                              n_old=n_old, iv_table=iv_table, prime_pad=self.prime_pad if self.training else 0,
                              path=self.compute_path)

    def _train_net(self, side):
        net, fm = self._side(side)
        return net, None, fm

    def _hash_users(self, users, feat_lookup_users=None):
        return self.assemble_rows("user", users, None, 0, None)

    def _hash_items(self, items, feat_lookup_items=None):
        return self.assemble_rows("item", items, None, 0, None)

    def embed_user_ids(self, old_user_ids, model) -> torch.Tensor:
        return self.assemble_rows("user", old_user_ids, model, 0, None)

    def embed_item_ids(self, old_item_ids, model) -> torch.Tensor:
        return self.assemble_rows("item", old_item_ids, model, 0, None)

    def embed_all_items(self, item_embeddings, model):
        raise NotImplementedError()
