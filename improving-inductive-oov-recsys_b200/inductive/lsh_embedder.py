"""`lsh` embedder — mirrors reference inductive/lsh_embedder.py:11-192 (same constructor, attributes,
state_dict keys `user_lsh.uniform_planes.0` / `item_lsh.uniform_planes.0`).

hash_size = n_oov_buckets planes; the hash is a multi-hot selector over the B OOV buckets and the
embedding is the mean of the selected bucket rows (lsh_embedder.py:156-158).  The GPU path packs the
selector into words with a warp ballot and never materialises the [n, B] fp32 matrix.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import ops
from .abstract_embedder import AbstractInductiveEmbedder, feature_block, feature_columns
from .torch_hash import TorchLSHash


def build_feature_mats(emb: AbstractInductiveEmbedder, normalization_type: str, device):
    """lsh_embedder.py:77-106 / single_lsh_embedder.py:56-75: per-feature / global / none."""
    user_columns = feature_columns(emb.user_features)[1:]
    item_columns = feature_columns(emb.item_features)[1:]
    if normalization_type == "per-feature":
        ufm = torch.hstack([F.normalize(feature_block(emb.user_features, c, emb.n_new_users), dim=-1) for c in user_columns])
        ifm = torch.hstack([F.normalize(feature_block(emb.item_features, c, emb.n_new_items), dim=-1) for c in item_columns])
    elif normalization_type in ("global", "none"):
        ufm = torch.hstack([feature_block(emb.user_features, c, emb.n_new_users) for c in user_columns])
        ifm = torch.hstack([feature_block(emb.item_features, c, emb.n_new_items) for c in item_columns])
    else:
        raise ValueError(f"Invalid normalization type: {normalization_type}")
    return ufm.to(device).contiguous(), ifm.to(device).contiguous()


class LSHInductiveEmbedder(AbstractInductiveEmbedder):
    def __init__(self, user_features, item_features, n_original_users, n_original_items, n_user_oov_buckets,
                 n_item_oov_buckets, embedding_size, device, prime_pad, normalization_type, feature_cache) -> None:
        super().__init__(user_features, item_features)
        self.n_original_users = n_original_users
        self.n_original_items = n_original_items
        self.n_user_oov_buckets = n_user_oov_buckets
        self.n_item_oov_buckets = n_item_oov_buckets
        self.embedding_size = embedding_size
        self.device = device
        self.prime_pad = prime_pad

        if feature_cache is not None and feature_cache.has_cached():
            self.user_feature_mat, self.item_feature_mat = feature_cache.get_cached()
        else:
            self.user_feature_mat, self.item_feature_mat = build_feature_mats(self, normalization_type, device)
            if normalization_type == "global":
                self.user_feature_mat = F.normalize(self.user_feature_mat, dim=-1)
                self.item_feature_mat = F.normalize(self.item_feature_mat, dim=-1)
            if feature_cache is not None:
                feature_cache.add_to_cache(self.user_feature_mat, self.item_feature_mat)

        self.user_lsh = TorchLSHash(hash_size=n_user_oov_buckets, input_dim=self.user_feature_mat.size(1), device=device)
        self.item_lsh = TorchLSHash(hash_size=n_item_oov_buckets, input_dim=self.item_feature_mat.size(1), device=device)
        self.tie_count = None        # optional int64[1] device counter of |x| < 1e-6 projections

    # --- hashing (lsh_embedder.py:116-139) -----------------------------------------------
    def _side(self, side: str):
        if side == "user":
            return self.user_lsh, self.user_feature_mat
        return self.item_lsh, self.item_feature_mat

    def _hash_node_packed(self, nodes: torch.Tensor, lsh: TorchLSHash, feature_mat: torch.Tensor) -> torch.Tensor:
        return ops.lsh_bits(feature_mat, lsh.uniform_planes[0].data, nodes, tie_count=self.tie_count)

    def _hash_node(self, nodes, lsh, feature_mat) -> torch.Tensor:
        words = self._hash_node_packed(nodes, lsh, feature_mat)
        shifts = torch.arange(32, device=words.device, dtype=torch.int32)
        bits = ((words.unsqueeze(-1) >> shifts) & 1).reshape(words.shape[0], -1)
        return bits[:, : lsh.hash_size].to(torch.float32)

    def _hash_users(self, users):
        return self._hash_node(users, self.user_lsh, self.user_feature_mat)

    def _hash_items(self, items):
        return self._hash_node(items, self.item_lsh, self.item_feature_mat)

    # --- embedding (lsh_embedder.py:141-179) ----------------------------------------------
    SIDE_CAST = True     # assemble_rows(side_cast=(src, dst)) folds a contiguous fp32 -> bf16 row cast into the launch

    def assemble_rows(self, side, ids, model, n_old, iv_table, out=None, out_dtype=torch.float32, side_cast=None):
        lsh, fm = self._side(side)
        w = (model.user_oov_buckets if side == "user" else model.item_oov_buckets).weight.detach()
        return ops.lsh_embed(fm, lsh.uniform_planes[0].data, w, ids, out=out, out_dtype=out_dtype, n_old=n_old,
                             iv_table=iv_table, prime_pad=self.prime_pad if self.training else 0,
                             tie_count=self.tie_count, side_cast=side_cast)

    # training: out = (H W) / |H| is linear in W = model.*_oov_buckets.weight; the backward re-uses the forward's bits
    def train_params(self, side, model):
        return [(model.user_oov_buckets if side == "user" else model.item_oov_buckets).weight]

    def assemble_rows_train(self, side, ids, model, n_old, iv_table):
        lsh, fm = self._side(side)
        w = self.train_params(side, model)[0].detach()
        out, bits = ops.lsh_embed(fm, lsh.uniform_planes[0].data, w, ids, n_old=n_old, iv_table=iv_table,
                                  prime_pad=self.prime_pad if self.training else 0, tie_count=self.tie_count, return_bits=True)
        return out, bits

    def backward_rows(self, side, saved, g, ids, n_old, model):
        w = self.train_params(side, model)[0]
        return [ops.lsh_embed_backward(saved, g, ids, n_old, torch.zeros_like(w, dtype=torch.float32))]

    def embed_user_ids(self, user_ids, model) -> torch.Tensor:
        self._depad_inplace(user_ids, self.prime_pad)
        return self.assemble_rows("user", user_ids, model, 0, None)

    def embed_item_ids(self, item_ids, model) -> torch.Tensor:
        self._depad_inplace(item_ids, self.prime_pad)
        return self.assemble_rows("item", item_ids, model, 0, None)

    def embed_all_items(self, item_embeddings, model):
        raise NotImplementedError()
