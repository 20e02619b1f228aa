"""Seeded synthetic inputs shared by the golden generator and the parity tests.

Pure numpy (PCG64 streams are stable across numpy versions), no reference and no
oracle imports: both `make_golden.py` (which feeds these inputs to the unmodified
reference) and the tests (which feed them to the oracle and to the CUDA path)
rebuild identical inputs from the seeds, so the committed fixtures only need to
hold the reference's OUTPUTS.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

OOV_PRIME_PAD = 112062759511  # RecBole/recbole/properties/overall.yaml:71


def rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def xavier_normal(g: np.random.Generator, rows: int, cols: int) -> np.ndarray:
    """nn.init.xavier_normal_ for an Embedding [rows, cols] (model/init.py: std = sqrt(2/(fan_in+fan_out)))."""
    std = np.sqrt(2.0 / (rows + cols))
    return (g.standard_normal((rows, cols)) * std).astype(np.float32)


def feature_columns(g: np.random.Generator, n: int, spec: List[Tuple[str, int]]) -> List[np.ndarray]:
    """spec: list of (kind, width).  kind 'float' -> fp32 [n, width] randn (row 0 = column
    mean, like the [PAD] entity, data/dataset/dataset.py:581-596); 'token' -> int64 [n]
    categorical ids (row 0 = 0); 'token_seq' -> int64 [n, width] zero-padded."""
    cols = []
    for kind, width in spec:
        if kind == "float":
            c = g.standard_normal((n, width)).astype(np.float32)
            c[0] = c[1:].mean(axis=0)
            if width == 1:
                c = c.reshape(n)
        elif kind == "token":
            c = g.integers(1, 20, size=n).astype(np.int64)
            c[0] = 0
        elif kind == "token_seq":
            c = g.integers(1, 50, size=(n, width)).astype(np.int64)
            lens = g.integers(1, width + 1, size=n)
            c[np.arange(width)[None, :] >= lens[:, None]] = 0
            c[0] = 0
        else:
            raise ValueError(kind)
        cols.append(c)
    return cols


@dataclass
class RetrievalCase:
    """One BPR/DirectAU + {lsh, slsh, mean, zero} retrieval case (SURVEY §8d config 1 shape)."""
    name: str
    seed: int
    embedder: str            # lsh | slsh | mean | zero | dhe
    model: str               # BPR | DirectAU
    n_old_users: int
    n_all_users: int         # rows of the inductive user feature table (old + new)
    n_old_items: int
    n_all_items: int
    D: int
    B_user: int
    B_item: int
    user_spec: List[Tuple[str, int]]
    item_spec: List[Tuple[str, int]]
    normalization: str = "per-feature"
    Q: int = 48
    k: int = 20
    max_hist: int = 30


# ml-100k-shaped: 944 users / 1683 items incl. pad row, +25 % OOV rows
CASES: Dict[str, RetrievalCase] = {
    "bpr_lsh_ml100k": RetrievalCase(
        "bpr_lsh_ml100k", 2020, "lsh", "BPR", 944, 1180, 1683, 2104, 64, 1000, 1000,
        [("token", 1), ("float", 8), ("float", 3), ("token", 1)],
        [("float", 8), ("float", 8), ("token_seq", 6), ("float", 1), ("token", 1)]),
    "directau_lsh_global": RetrievalCase(
        "directau_lsh_global", 7, "lsh", "DirectAU", 300, 420, 500, 777, 32, 64, 100,
        [("float", 5), ("token", 1)], [("float", 16), ("float", 7)], normalization="global", Q=33),
    "bpr_lsh_tinybuckets": RetrievalCase(   # small B: all-zero multi-hot rows -> NaN rows (lsh_embedder.py:158)
        "bpr_lsh_tinybuckets", 11, "lsh", "BPR", 100, 400, 120, 600, 16, 3, 2,
        [("float", 4)], [("float", 6)], normalization="none", Q=17, k=10),
    "directau_slsh": RetrievalCase(
        "directau_slsh", 2021, "slsh", "DirectAU", 944, 1180, 1683, 2104, 64, 1000, 1000,
        [("float", 8), ("float", 4)], [("float", 8), ("float", 8), ("float", 8), ("float", 8)]),
    "bpr_slsh_odd": RetrievalCase(
        "bpr_slsh_odd", 5, "slsh", "BPR", 50, 90, 70, 200, 24, 7, 33,
        [("float", 3)], [("float", 5), ("token", 1)], normalization="none", Q=9, k=5),
    "bpr_mean": RetrievalCase(
        "bpr_mean", 3, "mean", "BPR", 200, 260, 300, 450, 64, 10, 10,
        [("float", 4)], [("float", 4)], Q=21),
    "directau_zero": RetrievalCase(
        "directau_zero", 4, "zero", "DirectAU", 200, 260, 300, 450, 48, 10, 10,
        [("float", 4)], [("float", 4)], Q=21),
}


def retrieval_inputs(case: RetrievalCase) -> dict:
    g = rng(case.seed)
    out = {}
    out["user_cols"] = feature_columns(g, case.n_all_users, case.user_spec)
    out["item_cols"] = feature_columns(g, case.n_all_items, case.item_spec)
    out["user_table"] = xavier_normal(g, case.n_old_users, case.D)
    out["item_table"] = xavier_normal(g, case.n_old_items, case.D)
    out["user_oov"] = xavier_normal(g, case.B_user, case.D)
    out["item_oov"] = xavier_normal(g, case.B_item, case.D)
    f_user = sum(w for _, w in case.user_spec)
    f_item = sum(w for _, w in case.item_spec)
    if case.embedder == "lsh":
        pu, pi = case.B_user, case.B_item
    elif case.embedder == "slsh":
        pu = int(np.ceil(np.log2(case.B_user)))
        pi = int(np.ceil(np.log2(case.B_item)))
    else:
        pu = pi = 1
    out["user_planes"] = g.standard_normal((pu, f_user)).astype(np.float32)
    out["item_planes"] = g.standard_normal((pi, f_item)).astype(np.float32)
    # query users: half in-vocab, half OOV (ids >= n_old_users), shuffled
    q_iv = g.choice(np.arange(1, case.n_old_users), size=case.Q // 2, replace=False)
    q_oov = g.choice(np.arange(case.n_old_users, case.n_all_users), size=case.Q - case.Q // 2, replace=False)
    users = np.concatenate([q_iv, q_oov]).astype(np.int64)
    g.shuffle(users)
    out["users"] = users
    # history: 0..max_hist items per query row (evaluator.py:93-94 history_index = (row, item))
    hu, hi = [], []
    for r in range(case.Q):
        cnt = int(g.integers(0, case.max_hist + 1))
        if cnt:
            items = g.choice(np.arange(1, case.n_all_items), size=cnt, replace=False)
            hu.append(np.full(cnt, r, dtype=np.int64))
            hi.append(np.sort(items).astype(np.int64))
    out["hist_u"] = np.concatenate(hu) if hu else np.zeros(0, np.int64)
    out["hist_i"] = np.concatenate(hi) if hi else np.zeros(0, np.int64)
    # positives for the collector ('rec.topk'): 1..5 per row, disjoint from history
    pu_, pi_ = [], []
    for r in range(case.Q):
        cnt = int(g.integers(1, 6))
        banned = set(out["hist_i"][out["hist_u"] == r].tolist())
        cand = [x for x in g.choice(np.arange(1, case.n_all_items), size=cnt + len(banned), replace=False).tolist()
                if x not in banned][:cnt]
        pu_.append(np.full(len(cand), r, dtype=np.int64))
        pi_.append(np.asarray(sorted(cand), dtype=np.int64))
    out["pos_u"] = np.concatenate(pu_)
    out["pos_i"] = np.concatenate(pi_)
    return out


# ---------------------------------------------------------------------------------------
# DHE
# ---------------------------------------------------------------------------------------
@dataclass
class DheCase:
    name: str
    seed: int
    n_hashes: int = 128
    hidden: int = 512          # hard-coded in dh_embedder.py:70-89
    D: int = 64
    n_ids: int = 384
    w1_scale: float = 1.0      # 'trained-looking' variant scales layer 1 so activations are O(1)


DHE_CASES: Dict[str, DheCase] = {
    "dhe_default_init": DheCase("dhe_default_init", 101),
    "dhe_scaled": DheCase("dhe_scaled", 102, w1_scale=2e-7),
    "dhe_scaled_d16_h32": DheCase("dhe_scaled_d16_h32", 103, n_hashes=32, D=16, n_ids=200, w1_scale=4e-7),
}


def dhe_keys(seed: int, n_hashes: int) -> List[bytes]:
    g = rng(seed ^ 0x5EED)
    raw = g.integers(0, 256, size=(n_hashes, 16), dtype=np.uint8)
    return [bytes(r.tolist()) for r in raw]


def dhe_ids(case: DheCase) -> np.ndarray:
    g = rng(case.seed + 1)
    a = np.arange(0, case.n_ids // 2, dtype=np.int64)                         # small consecutive ids
    b = g.integers(0, 1 << 40, size=case.n_ids // 4, dtype=np.int64)          # large ids
    c = g.integers(0, 100000, size=case.n_ids - a.size - b.size, dtype=np.int64) + OOV_PRIME_PAD  # padded ids
    edge = np.array([0, 1, 255, 256, (1 << 31) - 1, 1 << 31, (1 << 32) - 1, 1 << 32, (1 << 63) - 1], dtype=np.int64)
    return np.concatenate([a, b, c, edge])


def dhe_weights(case: DheCase):
    """nn.Linear default init (kaiming_uniform a=sqrt(5) -> U(-1/sqrt(in), 1/sqrt(in)) for W and b)."""
    g = rng(case.seed + 2)
    dims = [case.n_hashes, case.hidden, case.hidden, case.hidden, case.D]
    ws, bs = [], []
    for l in range(4):
        bound = 1.0 / np.sqrt(dims[l])
        w = g.uniform(-bound, bound, size=(dims[l + 1], dims[l])).astype(np.float32)
        b = g.uniform(-bound, bound, size=(dims[l + 1],)).astype(np.float32)
        if l == 0:
            w = (w * np.float32(case.w1_scale)).astype(np.float32)
        ws.append(w)
        bs.append(b)
    return ws, bs


# ---------------------------------------------------------------------------------------
# Context (DCNV2 / WideDeep / xDeepFM) token gather + OOV overwrite
# ---------------------------------------------------------------------------------------
@dataclass
class ContextCase:
    name: str
    seed: int
    embedder: str
    batch: int
    n_fields: int
    D: int
    n_old_users: int
    n_all_users: int
    n_old_items: int
    n_all_items: int
    B: int
    F: int = 12
    oov_frac: float = 0.2


CONTEXT_CASES: Dict[str, ContextCase] = {
    "ctx_slsh": ContextCase("ctx_slsh", 31, "slsh", 512, 8, 16, 300, 400, 500, 650, 1000),
    "ctx_lsh": ContextCase("ctx_lsh", 32, "lsh", 300, 5, 10, 100, 160, 120, 200, 24),
    "ctx_mean": ContextCase("ctx_mean", 33, "mean", 512, 8, 10, 300, 400, 500, 650, 10),
    "ctx_zero": ContextCase("ctx_zero", 34, "zero", 257, 3, 10, 300, 400, 500, 650, 10),
    "ctx_no_oov": ContextCase("ctx_no_oov", 35, "slsh", 128, 4, 16, 300, 400, 500, 650, 16, oov_frac=0.0),
}


def context_inputs(case: ContextCase) -> dict:
    g = rng(case.seed)
    # field 0 = user id vocabulary, field 1 = item id vocabulary, rest categorical
    dims = [case.n_old_users, case.n_old_items] + [int(x) for x in g.integers(5, 2000, size=case.n_fields - 2)]
    offsets = np.concatenate([[0], np.cumsum(dims)[:-1]]).astype(np.int64)
    table = (g.standard_normal((int(sum(dims)), case.D)) * 0.1).astype(np.float32)
    table1 = (g.standard_normal((int(sum(dims)), 1)) * 0.1).astype(np.float32)   # first-order (D = 1) table
    tok = np.stack([g.integers(0, d, size=case.batch) for d in dims], axis=1).astype(np.int64)
    oov_u = g.random(case.batch) < case.oov_frac
    oov_i = g.random(case.batch) < case.oov_frac
    tok[oov_u, 0] = g.integers(case.n_old_users, case.n_all_users, size=int(oov_u.sum()))
    tok[oov_i, 1] = g.integers(case.n_old_items, case.n_all_items, size=int(oov_i.sum()))
    out = dict(dims=np.asarray(dims, np.int64), offsets=offsets, table=table, table1=table1, tokens=tok)
    out["user_cols"] = feature_columns(g, case.n_all_users, [("float", case.F)])
    out["item_cols"] = feature_columns(g, case.n_all_items, [("float", case.F // 2), ("float", case.F - case.F // 2)])
    out["user_oov"] = xavier_normal(g, case.B, case.D)
    out["item_oov"] = xavier_normal(g, case.B, case.D)
    out["user_oov1"] = xavier_normal(g, case.B, 1)
    out["item_oov1"] = xavier_normal(g, case.B, 1)
    if case.embedder == "lsh":
        p = case.B
    elif case.embedder == "slsh":
        p = int(np.ceil(np.log2(case.B)))
    else:
        p = 1
    for nm in ("user_planes", "item_planes", "user_planes1", "item_planes1"):   # *1 = first-order embedder's own planes
        out[nm] = g.standard_normal((p, case.F)).astype(np.float32)
    return out


# ---------------------------------------------------------------------------------------
# fdhe / dnn (feat_dh_embedder.py:86-210, dnn_embedder.py:8-112): the DHE net fed with [hashes | feature row]
# ---------------------------------------------------------------------------------------
@dataclass
class FeatNetCase:
    name: str
    seed: int
    kind: str                  # fdhe | dnn
    n_hashes: int = 128        # fdhe only
    layer: int = 512           # dhe_layer_size
    D: int = 64
    n_all: int = 300           # rows of the feature frames (every id the embedder may see)
    n_ids: int = 257
    h_scale: float = 2e-7      # 'trained-looking' scale of the hash columns of layer 1 (raw hashes are ~1e7)
    widths: Tuple[int, ...] = (4, 3)   # float feature columns -> F = sum(widths)


FEATNET_CASES: Dict[str, FeatNetCase] = {
    "fdhe_f7": FeatNetCase("fdhe_f7", 201, "fdhe"),
    "fdhe_small": FeatNetCase("fdhe_small", 202, "fdhe", n_hashes=32, layer=128, D=16, n_ids=131, h_scale=4e-7, widths=(8, 1, 16)),
    "dnn_f7": FeatNetCase("dnn_f7", 203, "dnn"),
    "dnn_wide": FeatNetCase("dnn_wide", 204, "dnn", layer=256, D=32, n_ids=200, widths=(24, 8, 5)),
}


def featnet_inputs(case: FeatNetCase) -> dict:
    g = rng(case.seed)
    ucols = feature_columns(g, case.n_all, [("float", w) for w in case.widths])
    icols = feature_columns(g, case.n_all, [("float", w) for w in case.widths])
    F = int(sum(case.widths)) if case.kind != "dhe" else 0          # the plain dhe net takes hashes only (dh_embedder.py:71)
    H = case.n_hashes if case.kind in ("fdhe", "dhe") else 0
    dims = [H + F, case.layer, case.layer, case.layer, case.D]
    nets = {}
    for side in ("user", "item"):
        ws, bs = [], []
        for l in range(4):
            bound = 1.0 / np.sqrt(dims[l])
            w = g.uniform(-bound, bound, size=(dims[l + 1], dims[l])).astype(np.float32)
            b = g.uniform(-bound, bound, size=(dims[l + 1],)).astype(np.float32)
            if l == 0 and H:
                w[:, :H] *= np.float32(case.h_scale)
                if F:
                    w[:, H:] *= np.float32(np.sqrt(dims[0] / F))       # features carry weight next to the hashes
            ws.append(w)
            bs.append(b)
        nets[side] = (ws, bs)
    ids = g.integers(0, case.n_all, size=case.n_ids, dtype=np.int64)
    ids[:4] = [0, 1, case.n_all - 1, case.n_all // 2]
    pad = g.random(case.n_ids) < 0.5                                   # training mode: half of the ids carry the prime pad
    ids_train = ids + pad.astype(np.int64) * OOV_PRIME_PAD
    return dict(user_cols=ucols, item_cols=icols, nets=nets, ids=ids, ids_train=ids_train, F=F, H=H)


# ---------------------------------------------------------------------------------------
# Training-mode OOV batches (trainer.py:1748-1837 _train_oov + _transform_interaction_oov)
# ---------------------------------------------------------------------------------------
TRAIN_CASES = ["bpr_lsh_ml100k", "directau_lsh_global", "bpr_lsh_tinybuckets", "directau_slsh", "bpr_slsh_odd", "directau_zero", "bpr_mean"]


def train_batch(case: RetrievalCase, batch: int = 192) -> dict:
    """A training batch of in-vocab (user, pos item, neg item) ids; a seeded per-element mask adds the prime pad to user
    and item ids (the reference pads whole columns by a coin flip, trainer.py:1748-1753 — a per-element mask covers both
    outcomes and mixes in-vocab and OOV rows in one batch).  Ids repeat, so the scatter-adds collide."""
    g = rng(case.seed + 77)
    users = g.integers(1, case.n_old_users, size=batch, dtype=np.int64)
    pos = g.integers(1, case.n_old_items, size=batch, dtype=np.int64)
    neg = g.integers(1, case.n_old_items, size=batch, dtype=np.int64)
    users[: batch // 8] = users[0]                       # heavy collisions on one row
    pad_u = g.random(batch) < 0.5
    pad_i = g.random(batch) < 0.5
    return dict(users=users + pad_u * OOV_PRIME_PAD, pos=pos + pad_i * OOV_PRIME_PAD, neg=neg)


# hash-net embedders under training (BPR on top; n_old = n_all // 2 in-vocab ids per side)
HASHNET_TRAIN_CASES: Dict[str, FeatNetCase] = {
    "train_fdhe": FeatNetCase("train_fdhe", 211, "fdhe", n_hashes=32, layer=64, D=16, n_all=240, h_scale=4e-7, widths=(8, 1, 6)),
    "train_dnn": FeatNetCase("train_dnn", 212, "dnn", layer=96, D=24, n_all=200, widths=(5, 8)),
    "train_dhe": FeatNetCase("train_dhe", 213, "dhe", n_hashes=16, layer=512, D=16, n_all=200, h_scale=8e-7),
}


def hashnet_train_inputs(case: FeatNetCase, batch: int = 160) -> dict:
    inp = featnet_inputs(case)
    g = rng(case.seed + 99)
    n_old = case.n_all // 2
    inp["n_old"] = n_old
    inp["user_table"] = xavier_normal(g, n_old, case.D)
    inp["item_table"] = xavier_normal(g, n_old, case.D)
    users = g.integers(1, n_old, size=batch, dtype=np.int64)
    pos = g.integers(1, n_old, size=batch, dtype=np.int64)
    neg = g.integers(1, n_old, size=batch, dtype=np.int64)
    inp["batch"] = dict(users=users + (g.random(batch) < 0.5) * OOV_PRIME_PAD, pos=pos + (g.random(batch) < 0.5) * OOV_PRIME_PAD, neg=neg)
    return inp


def grad_slice(a: np.ndarray) -> np.ndarray:
    """Fixtures keep every row of a small gradient and every s-th row of a large one (<= ~20k elements)."""
    a = np.asarray(a)
    if a.ndim < 2 or a.size <= 20000:
        return a
    return a[:: int(np.ceil(a.size / 20000))]
