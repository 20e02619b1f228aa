"""`dhe` embedder — mirrors reference inductive/dh_embedder.py:18-259.

Same constructor, `HASH_KEY_PATH` / `MAX_HASH`, key-file protocol (`./hash_keys/{num_hashes}.hashes`,
JSON list of hex strings, dh_embedder.py:95-120) and state_dict keys
(`user_hash_net.{0,2,4,6}.{weight,bias}`, `item_hash_net...`).  The reference hashes each id in a
Python loop through the csiphash C wheel (dh_embedder.py:154-170); here one CUDA thread computes one
(id, key) SipHash-2-4 and the 4-layer net runs on the GPU.
"""
from __future__ import annotations

import json
import os
import secrets

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .abstract_embedder import AbstractInductiveEmbedder, feature_block, feature_columns


class HashNetTraining:
    """Training-mode path shared by the dhe / fdhe / dnn embedders: the 4-layer net in fp32 with the
    pre-activations kept, and its backward (dZ = dA * act'(Z), dW = dZ^T A_prev, db = colsum dZ, dA_prev = dZ W) on the
    fp32 CUDA-core linear.  Embedders provide `_train_net(side)` -> (nn.Sequential, keys or None, feature matrix or None)."""

    _ACTS = ("gelu", "gelu", "gelu", "sigmoid")

    def train_params(self, side, model):
        net = self._train_net(side)[0]
        lin = [m for m in net if isinstance(m, nn.Linear)]
        return [p for m in lin for p in (m.weight, m.bias)]

    def assemble_rows_train(self, side, ids, model, n_old, iv_table):
        net, keys, fm = self._train_net(side)
        lin = [m for m in net if isinstance(m, nn.Linear)]
        a = ops.fdhe_input(ids, keys, fm, prime_pad=self.prime_pad if (self.training and fm is not None) else 0)
        acts, pre = [a], []
        for m, act in zip(lin, self._ACTS):
            z = ops.linear_f32(a, m.weight.detach(), m.bias.detach(), act="none")
            a = ops.act_forward(z, act)
            pre.append(z)
            acts.append(a)
        out = a
        if iv_table is not None and n_old > 0:
            ops.gather_rows(iv_table, ids, out=out)                # in-vocab rows take the table row (ids >= n_old are skipped)
        return out, (acts[:-1], pre)

    def backward_rows(self, side, saved, g, ids, n_old, model):
        net = self._train_net(side)[0]
        lin = [m for m in net if isinstance(m, nn.Linear)]
        acts, pre = saved
        grads = [None] * 8
        da = g
        for l in (3, 2, 1, 0):
            dz = ops.act_backward(pre[l], da, self._ACTS[l], ids if (l == 3 and n_old > 0) else None, n_old)
            grads[2 * l] = ops.linear_f32(dz.t().contiguous(), acts[l].t().contiguous())          # dW = dZ^T A_prev  [out, in]
            grads[2 * l + 1] = ops.col_mean(dz) * dz.shape[0]                                      # db = colsum dZ
            if l:
                da = ops.linear_f32(dz, lin[l].weight.detach().t().contiguous())                    # dA_prev = dZ W
        return grads


class DeepHashEmbedder(HashNetTraining, AbstractInductiveEmbedder):
    HASH_KEY_PATH = "./hash_keys"
    MAX_HASH = 16777216

    def __init__(self, user_features, item_features, n_original_users, n_original_items, n_user_oov_buckets,
                 n_item_oov_buckets, embedding_size, device, prime_pad, num_hashes) -> None:
        super().__init__(user_features, item_features)
        self.n_original_users = n_original_users
        self.n_original_items = n_original_items
        self.n_user_oov_buckets = n_user_oov_buckets
        self.n_item_oov_buckets = n_item_oov_buckets
        self.embedding_size = embedding_size
        self.device = device
        self.prime_pad = prime_pad
        self.num_hashes = num_hashes

        def hash_net():
            return nn.Sequential(
                nn.Linear(self.num_hashes, 512), nn.GELU(),
                nn.Linear(512, 512), nn.GELU(),
                nn.Linear(512, 512), nn.GELU(),
                nn.Linear(512, self.embedding_size), nn.Sigmoid()).to(self.device)

        self.user_hash_net = hash_net()
        self.item_hash_net = hash_net()

        # built but unused by DHE, exactly like dh_embedder.py:91-92 (read by abstract_recommender.py:751-753)
        user_columns = feature_columns(self.user_features)[1:]
        item_columns = feature_columns(self.item_features)[1:]
        self.user_feature_mat = torch.hstack(
            [F.normalize(feature_block(self.user_features, c, self.n_new_users), dim=-1) for c in user_columns]).to(device)
        self.item_feature_mat = torch.hstack(
            [F.normalize(feature_block(self.item_features, c, self.n_new_items), dim=-1) for c in item_columns]).to(device)

        self.hash_keys = self.get_hash_keys()
        self._keys_dev = ops.keys_tensor(self.hash_keys, self.device)
        self.compute_path = ops.PATH_AUTO
        self._planes_cache = {}            # (lo, hi, device) -> bf16 byte planes of arange(lo, hi)

    def get_hash_keys(self):
        os.makedirs(DeepHashEmbedder.HASH_KEY_PATH, exist_ok=True)
        file_path = os.path.join(DeepHashEmbedder.HASH_KEY_PATH, f"{self.num_hashes}.hashes")
        if os.path.exists(file_path):
            with open(file_path) as f:
                keys = json.load(f)
                assert len(keys) == self.num_hashes
                return [bytes.fromhex(x) for x in keys]
        keys = [secrets.token_bytes(16) for _ in range(self.num_hashes)]
        with open(file_path, "w") as f:
            json.dump([x.hex() for x in keys], f)
        return keys

    # --- hashing (dh_embedder.py:122-170) -------------------------------------------------
    def _hash_ids(self, ids: torch.Tensor) -> torch.Tensor:
        """fp32 [n, num_hashes] of exact integers < 2^24 (dh_embedder.py:154-170)."""
        return ops.dhe_hash(ids.to(self.device), self._keys_dev, DeepHashEmbedder.MAX_HASH).to(torch.float32)

    def _hash_id(self, id: torch.Tensor) -> torch.Tensor:
        return self._hash_ids(id.reshape(1))[0].to(torch.double)

    def _net(self, side: str) -> ops.DheNet:
        return ops.DheNet.from_sequential(self.user_hash_net if side == "user" else self.item_hash_net)

    def _hash_users(self, users):
        return self.assemble_rows("user", users, None, 0, None)

    def _hash_items(self, items):
        return self.assemble_rows("item", items, None, 0, None)

    def assemble_rows(self, side, ids, model, n_old, iv_table, out=None, out_dtype=torch.float32, id_range=None):
        """`id_range=(lo, hi)`: the caller states that `ids` is arange(lo, hi) (the all-item table, a row shard of it).
        The hashes of an id depend on the id and the keys only — the reference memoises them per id (dh_embedder.py:139
        `@cache` on _get_hashes) — so the byte planes of such a range are computed once and kept (768 B per id); later
        calls run the MLP alone.  Arbitrary id tensors (query batches) are hashed on every call."""
        net = self._net(side)
        use_tc = out_dtype == torch.bfloat16 or self.compute_path == ops.PATH_TCGEN05
        if id_range is not None and use_tc and self.compute_path != ops.PATH_SIMT_FP32 and net.hidden % 8 == 0 and net.D >= 8:
            key = (int(id_range[0]), int(id_range[1]), ids.device.index)
            planes = self._planes_cache.get(key)
            if planes is None:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("DHE hash planes must be memoised before a CUDA graph is captured (run the step once eagerly)")
                planes = ops.dhe_hash_planes(ids, self._keys_dev, DeepHashEmbedder.MAX_HASH)
                while len(self._planes_cache) >= 4:                      # a few shards at most
                    self._planes_cache.pop(next(iter(self._planes_cache)))
                self._planes_cache[key] = planes
            return ops.dhe_embed_planes(planes, ids, net, out=out, out_dtype=out_dtype, n_old=n_old, iv_table=iv_table)
        return ops.dhe_embed(ids, self._keys_dev, net, out=out, out_dtype=out_dtype, n_old=n_old,
                             iv_table=iv_table, mod=DeepHashEmbedder.MAX_HASH, path=self.compute_path)

    def _train_net(self, side):
        return (self.user_hash_net if side == "user" else self.item_hash_net), self._keys_dev, None

    def clear_hash_cache(self):
        """Drop the memoised hash planes (call after changing `hash_keys` / `_keys_dev`)."""
        self._planes_cache.clear()

    def embed_user_ids(self, user_ids, model) -> torch.Tensor:
        return self._hash_users(user_ids)

    def embed_item_ids(self, item_ids, model) -> torch.Tensor:
        return self._hash_items(item_ids)

    def embed_all_items(self, item_embeddings, model):
        raise NotImplementedError()
