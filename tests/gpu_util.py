"""Build the PRODUCT objects (embedders, models) for a seeded case on cuda:0 — the GPU-side
counterpart of parity_util.oracle_retrieval.  Everything goes through the public plugin API
(get_inductive_embedder -> embedder -> model), i.e. through the C-ABI."""
from __future__ import annotations

import numpy as np
import torch

import cases
import oov_b200
from oov_b200 import ops

DEV = "cuda:0"


class Config(dict):
    """Missing keys read as None, like RecBole's Config (config/configurator.py:583-584)."""

    def __getitem__(self, key):
        return dict.get(self, key, None)


class Dataset:
    def __init__(self, n_users, n_items, user_feat, item_feat):
        self._num = {"user_id": n_users, "item_id": n_items}
        self.user_num, self.item_num = n_users, n_items
        self._uf, self._if = user_feat, item_feat

    def num(self, field):
        return self._num[field]

    def get_user_feature(self):
        return self._uf

    def get_item_feature(self):
        return self._if


def interaction(id_field, cols):
    d = {id_field: torch.arange(cols[0].shape[0])}
    for i, c in enumerate(cols):
        d[f"f{i}"] = torch.from_numpy(c)
    return oov_b200.Interaction(d)


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return x if dtype is None else x.to(dtype)


def make_config(case_like, embedder, **extra):
    cfg = Config(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device=DEV,
                 embedding_size=case_like.D, add_oov_buckets=True, inductive_embedder=embedder,
                 oov_prime_pad=cases.OOV_PRIME_PAD,
                 oov_normalization_type=getattr(case_like, "normalization", "per-feature"),
                 dhe_num_hashes=128, gamma=1.0, topk=[10, 20], model_eval_type="retrieval")
    cfg.update(extra)
    return cfg


def build_retrieval(case: cases.RetrievalCase, inp: dict, table_dtype="float32"):
    uf = interaction("user_id", inp["user_cols"])
    itf = interaction("item_id", inp["item_cols"])
    cfg = make_config(case, case.embedder, user_oov_buckets=case.B_user, item_oov_buckets=case.B_item,
                      table_dtype=table_dtype, topk=[min(10, case.k), case.k])
    ds = Dataset(case.n_old_users, case.n_old_items, uf, itf)
    emb = oov_b200.get_inductive_embedder(cfg, ds, mode=f"test-{case.name}", user_num=case.n_old_users,
                                          item_num=case.n_old_items)
    if case.embedder in ("lsh", "slsh"):
        emb.user_lsh.uniform_planes[0].data.copy_(t(inp["user_planes"]))
        emb.item_lsh.uniform_planes[0].data.copy_(t(inp["item_planes"]))
        emb.tie_count = ops.new_counter(DEV)
    cls = oov_b200.BPR if case.model == "BPR" else oov_b200.DirectAU
    model = cls(cfg, ds, inductive_mapper=None, inductive_embedder=emb).to(DEV).eval()
    with torch.no_grad():
        model.user_embedding.weight.copy_(t(inp["user_table"]))
        model.item_embedding.weight.copy_(t(inp["item_table"]))
        model.user_oov_buckets.weight.copy_(t(inp["user_oov"]))
        model.item_oov_buckets.weight.copy_(t(inp["item_oov"]))
    return cfg, emb, model
