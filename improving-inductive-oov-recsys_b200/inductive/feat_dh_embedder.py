"""`fdhe` embedder — mirrors reference inductive/feat_dh_embedder.py:86-210.

Same constructor (incl. `dhe_layer_size`), `HASH_KEY_PATH` / `MAX_HASH`, key-file protocol and state_dict keys
(`user_hash_net.{0,2,4,6}.{weight,bias}`, `item_hash_net...`) as the reference.  The first Linear takes
`hstack(hashes of the id, feature row of the id)` (feat_dh_embedder.py:188-196): here one call
(`oov_fdhe_embed`) hashes, fetches the feature row, runs the four layers and assembles in-vocab / OOV rows.
In training mode the hashes use the ORIGINAL (padded) id and only the feature lookup is de-padded
(feat_dh_embedder.py:198-213) — unlike `lsh`, the caller's ids are not modified.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .abstract_embedder import AbstractInductiveEmbedder, feature_block, feature_columns
from .dh_embedder import DeepHashEmbedder, HashNetTraining


def _hash_net(in_dim: int, layer: int, out_dim: int, device) -> nn.Sequential:
    return nn.Sequential(
        nn.Linear(in_dim, layer), nn.GELU(),
        nn.Linear(layer, layer), nn.GELU(),
        nn.Linear(layer, layer), nn.GELU(),
        nn.Linear(layer, out_dim), nn.Sigmoid()).to(device)


def _feature_mats(emb: AbstractInductiveEmbedder, device):
    """Per-column L2-normalised blocks, hstacked (feat_dh_embedder.py:96-99, dnn_embedder.py:63-64)."""
    user_columns = feature_columns(emb.user_features)[1:]
    item_columns = feature_columns(emb.item_features)[1:]
    um = torch.hstack([F.normalize(feature_block(emb.user_features, c, emb.n_new_users), dim=-1) for c in user_columns]).to(device)
    im = torch.hstack([F.normalize(feature_block(emb.item_features, c, emb.n_new_items), dim=-1) for c in item_columns]).to(device)
    return um.contiguous(), im.contiguous()


class FeatDeepHashEmbedder(HashNetTraining, AbstractInductiveEmbedder):
    HASH_KEY_PATH = "./hash_keys"
    MAX_HASH = 16777216

    def __init__(self, user_features, item_features, n_original_users, n_original_items, n_user_oov_buckets,
                 n_item_oov_buckets, embedding_size, device, prime_pad, num_hashes, dhe_layer_size) -> None:
        super().__init__(user_features, item_features)
        self.n_original_users = n_original_users
        self.n_original_items = n_original_items
        self.n_user_oov_buckets = n_user_oov_buckets
        self.n_item_oov_buckets = n_item_oov_buckets
        self.embedding_size = embedding_size
        self.device = device
        self.prime_pad = prime_pad
        self.num_hashes = num_hashes
        self.user_feature_mat, self.item_feature_mat = _feature_mats(self, device)
        self.user_hash_net = _hash_net(self.num_hashes + self.user_feature_mat.size(1), dhe_layer_size, embedding_size, device)
        self.item_hash_net = _hash_net(self.num_hashes + self.item_feature_mat.size(1), dhe_layer_size, embedding_size, device)
        self.hash_keys = self.get_hash_keys()
        self._keys_dev = ops.keys_tensor(self.hash_keys, self.device)
        self.compute_path = ops.PATH_AUTO

    get_hash_keys = DeepHashEmbedder.get_hash_keys            # same ./hash_keys/{n}.hashes protocol (feat_dh_embedder.py:131-150)

    def _hash_ids(self, ids: torch.Tensor) -> torch.Tensor:
        return ops.dhe_hash(ids.to(self.device), self._keys_dev, FeatDeepHashEmbedder.MAX_HASH).to(torch.float32)

    def _hash_id(self, id: torch.Tensor) -> torch.Tensor:
        return self._hash_ids(id.reshape(1))[0].to(torch.double)

    def _side(self, side: str):
        if side == "user":
            return self.user_hash_net, self.user_feature_mat
        return self.item_hash_net, self.item_feature_mat

    def assemble_rows(self, side, ids, model, n_old, iv_table, out=None, out_dtype=torch.float32):
        net, fm = self._side(side)
        return ops.fdhe_embed(ids, self._keys_dev, ops.DheNet.from_sequential(net, n_feat=fm.shape[1]), fm, out=out,
                              out_dtype=out_dtype, n_old=n_old, iv_table=iv_table,
                              prime_pad=self.prime_pad if self.training else 0, mod=FeatDeepHashEmbedder.MAX_HASH,
                              path=self.compute_path)

    def _train_net(self, side):
        net, fm = self._side(side)
        return net, self._keys_dev, fm

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
