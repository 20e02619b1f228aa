"""ORACLE — CPU restatement of the reference's OOV hot path (test infrastructure only).

This file is NOT product code.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product
package (`improving-inductive-oov-recsys_b200/`) never does and has no CPU fallback.

Every function restates one piece of `/root/reference` in plain numpy (fp32 where
the reference is fp32) and cites the file:line it follows.  Citations are relative
to `/root/reference/RecBole/recbole/` unless they start with `src/`.

Parity pin
----------
The reference has **no** golden vectors or tests for this path (no file under
`RecBole/tests/` mentions inductive/oov/lsh/dhe).  The oracle is therefore pinned
against outputs of the reference itself, run in the authoring container through
`oracle/refshim.py`: `tests/golden/make_golden.py` imports the unmodified reference
classes, runs them on seeded inputs and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function below against those fixtures
(and `tests/test_oracle_vs_reference.py` re-checks live when the reference tree is
present).  The one boundary that stays "parity unpinned" by the reference is the
third-party `csiphash==0.0.5` wheel (absent, un-vendored): SipHash-2-4 is restated
from its published specification in `oracle/siphash24.c` and pinned with the SipHash
paper's known-answer vectors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Callable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_siphash.so")
_lib = None

MAX_HASH = 16777216  # inductive/dh_embedder.py:53
OOV_PRIME_PAD = 112062759511  # properties/overall.yaml:71


# ---------------------------------------------------------------------------------------
# SipHash-2-4 (csiphash stand-in) — C restatement + a pure-Python cross-check
# ---------------------------------------------------------------------------------------
def build_c(force: bool = False) -> str:
    """Compile oracle/siphash24.c into oracle/liboracle_siphash.so (gcc, -O2)."""
    src = os.path.join(_HERE, "siphash24.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def _c():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_siphash24.restype = ctypes.c_uint64
        lib.oracle_siphash24.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        lib.oracle_dhe_hashes.restype = None
        lib.oracle_dhe_hashes.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p,
                                          ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]
        _lib = lib
    return _lib


def siphash24_u64(key: bytes, msg: bytes) -> int:
    assert len(key) == 16
    return int(_c().oracle_siphash24(key, msg, len(msg)))


def siphash24_bytes(key: bytes, msg: bytes) -> bytes:
    """csiphash.siphash24 convention: 8 little-endian bytes (dh_embedder.py:152)."""
    return siphash24_u64(key, msg).to_bytes(8, "little")


_M64 = (1 << 64) - 1


def siphash24_py(key: bytes, msg: bytes) -> int:
    """Pure-Python SipHash-2-4 (slow; cross-checks the C restatement on small cases)."""
    def rotl(x, b):
        return ((x << b) | (x >> (64 - b))) & _M64

    k0 = int.from_bytes(key[:8], "little")
    k1 = int.from_bytes(key[8:], "little")
    v = [k0 ^ 0x736F6D6570736575, k1 ^ 0x646F72616E646F6D,
         k0 ^ 0x6C7967656E657261, k1 ^ 0x7465646279746573]

    def rnd():
        v[0] = (v[0] + v[1]) & _M64; v[1] = rotl(v[1], 13); v[1] ^= v[0]; v[0] = rotl(v[0], 32)
        v[2] = (v[2] + v[3]) & _M64; v[3] = rotl(v[3], 16); v[3] ^= v[2]
        v[0] = (v[0] + v[3]) & _M64; v[3] = rotl(v[3], 21); v[3] ^= v[0]
        v[2] = (v[2] + v[1]) & _M64; v[1] = rotl(v[1], 17); v[1] ^= v[2]; v[2] = rotl(v[2], 32)

    n = len(msg)
    for i in range(n // 8):
        m = int.from_bytes(msg[8 * i:8 * i + 8], "little")
        v[3] ^= m; rnd(); rnd(); v[0] ^= m
    b = (n << 56) & _M64
    b |= int.from_bytes(msg[8 * (n // 8):], "little")
    v[3] ^= b; rnd(); rnd(); v[0] ^= b
    v[2] ^= 0xFF
    rnd(); rnd(); rnd(); rnd()
    return v[0] ^ v[1] ^ v[2] ^ v[3]


def keys_to_array(keys: Sequence[bytes]) -> np.ndarray:
    """[n_hashes] 16-byte keys -> uint8 [n_hashes, 16] (the layout the C-ABI takes)."""
    arr = np.frombuffer(b"".join(keys), dtype=np.uint8).reshape(len(keys), 16).copy()
    return arr


def dhe_hashes(ids: np.ndarray, keys: np.ndarray) -> np.ndarray:
    """dh_embedder.py:140-170: out[i,j] = LE_u64(siphash24(key_j, LE8(id_i))) % 2**24.

    Returns uint32 [n, n_hashes]; the reference then holds these as exact fp32 values.
    """
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    keys = np.ascontiguousarray(keys, dtype=np.uint8)
    out = np.empty((ids.shape[0], keys.shape[0]), dtype=np.uint32)
    _c().oracle_dhe_hashes(ids.ctypes.data, ids.shape[0], keys.ctypes.data_as(ctypes.c_char_p),
                           keys.shape[0], MAX_HASH, out.ctypes.data)
    return out


# ---------------------------------------------------------------------------------------
# Feature matrices (lsh_embedder.py:77-106, single_lsh_embedder.py:56-75)
# ---------------------------------------------------------------------------------------
def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """torch.nn.functional.normalize(x, dim=-1): x / max(||x||_2, eps), fp32."""
    x = np.asarray(x, dtype=np.float32)
    nrm = np.sqrt((x * x).sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return (x / np.maximum(nrm, np.float32(eps))).astype(np.float32)


def feature_matrix(columns: Sequence[np.ndarray], normalization_type: str = "per-feature") -> np.ndarray:
    """hstack of per-column blocks, each `.float().view(n, -1)`; normalised per block
    ('per-feature'), over the whole row ('global') or not at all ('none')."""
    n = columns[0].shape[0]
    blocks = [np.asarray(c).astype(np.float32).reshape(n, -1) for c in columns]
    if normalization_type == "per-feature":
        blocks = [l2_normalize(b) for b in blocks]
    elif normalization_type not in ("global", "none"):
        raise ValueError(f"Invalid normalization type: {normalization_type}")
    mat = np.hstack(blocks).astype(np.float32)
    if normalization_type == "global":
        mat = l2_normalize(mat)
    return mat


# ---------------------------------------------------------------------------------------
# LSH / SLSH (torch_hash.py:55-60, lsh_embedder.py:116-179, single_lsh_embedder.py:77-109)
# ---------------------------------------------------------------------------------------
def depad_ids(ids: np.ndarray, training: bool, prime_pad: int = OOV_PRIME_PAD) -> np.ndarray:
    """lsh_embedder.py:153-155 / 173-175: in training mode ids >= prime_pad lose the pad."""
    ids = np.asarray(ids, dtype=np.int64)
    if training:
        ids = np.where(ids >= prime_pad, ids - prime_pad, ids)
    return ids


def projections(planes: np.ndarray, points: np.ndarray) -> np.ndarray:
    """torch_hash.py:56: result = input_points @ planes.T  (fp32)."""
    return (np.asarray(points, np.float32) @ np.asarray(planes, np.float32).T).astype(np.float32)


def hash_points(planes: np.ndarray, points: np.ndarray) -> np.ndarray:
    """torch_hash.py:55-60: R < 0 -> 0, everything else (+0, -0, NaN, >0) -> 1; fp32 0/1."""
    r = projections(planes, points)
    return np.where(r < 0, np.float32(0), np.float32(1)).astype(np.float32)


def lsh_multihot(feature_mat, ids, planes, training=False, prime_pad=OOV_PRIME_PAD):
    """lsh_embedder.py:116-131: multi-hot selector [n, B] over the B OOV buckets."""
    ids = depad_ids(ids, training, prime_pad)
    return hash_points(planes, feature_mat[ids])


def lsh_embed(feature_mat, ids, planes, oov_weight, training=False, prime_pad=OOV_PRIME_PAD):
    """lsh_embedder.py:141-179: (H @ W) / H.sum(1)  — mean of the selected bucket rows.
    An all-zero multi-hot row gives 0/0 = NaN, as in the reference."""
    h = lsh_multihot(feature_mat, ids, planes, training, prime_pad)
    num = (h @ np.asarray(oov_weight, np.float32)).astype(np.float32)
    den = h.sum(axis=1, dtype=np.float32).reshape(-1, 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (num / den).astype(np.float32)


def slsh_bits_req(n_buckets: int) -> int:
    """single_lsh_embedder.py:77-78: int(ceil(log2(n_buckets)))."""
    return int(np.ceil(np.log2(n_buckets)))


def slsh_ids(feature_mat, ids, planes, n_buckets, training=False, prime_pad=OOV_PRIME_PAD):
    """single_lsh_embedder.py:82-87: ((2 ** H).sum(1)).long() % n_buckets, H in {0,1}.
    (= (bits_req + popcount(H)) % n_buckets — NOT a packed integer.)"""
    ids = depad_ids(ids, training, prime_pad)
    h = hash_points(planes, feature_mat[ids])
    return (np.power(np.float32(2), h).sum(axis=1).astype(np.int64)) % int(n_buckets)


def slsh_embed(feature_mat, ids, planes, oov_weight, training=False, prime_pad=OOV_PRIME_PAD):
    """single_lsh_embedder.py:95-109: model.*_oov_buckets(bucket_id) row lookup."""
    b = slsh_ids(feature_mat, ids, planes, oov_weight.shape[0], training, prime_pad)
    return np.asarray(oov_weight, np.float32)[b]


# ---------------------------------------------------------------------------------------
# DHE MLP (dh_embedder.py:70-89, 191-217)
# ---------------------------------------------------------------------------------------
def gelu_erf(x: np.ndarray) -> np.ndarray:
    """nn.GELU() default (approximate='none'): 0.5 * x * (1 + erf(x / sqrt(2)))."""
    from scipy.special import erf

    x = np.asarray(x, np.float32)
    return (np.float32(0.5) * x * (np.float32(1) + erf(x * np.float32(0.7071067811865476)).astype(np.float32))).astype(np.float32)


def sigmoid(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.float32)
    with np.errstate(over="ignore"):
        return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)


def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) -> fp32.  Used to emulate the
    bf16 rounding points of the tensor-core path so that only accumulation order
    differs between the kernel and this restatement."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    nan = np.isnan(x)
    out[nan] = np.nan
    return out


def dhe_mlp(hashes: np.ndarray, weights: Sequence[np.ndarray], biases: Sequence[np.ndarray],
            bf16_points: bool = False) -> np.ndarray:
    """dh_embedder.py:70-89: Linear-GELU ×3, Linear-Sigmoid on the raw fp32 hash values.

    weights[l] is [out, in] like nn.Linear.weight.  With `bf16_points=True` the
    hidden activations are rounded to bf16 after each GELU (the tensor-core path's
    operand precision); layer-1 inputs stay exact 24-bit integers either way.
    """
    x = np.asarray(hashes).astype(np.float32)
    n_layers = len(weights)
    for l in range(n_layers):
        w = np.asarray(weights[l], np.float32)
        b = np.asarray(biases[l], np.float32)
        # fp64 accumulate then round: the order-independent stand-in for an fp32 GEMM
        x = (x.astype(np.float64) @ w.T.astype(np.float64) + b.astype(np.float64)).astype(np.float32)
        if l < n_layers - 1:
            x = gelu_erf(x)
            if bf16_points:
                x = round_bf16(x)
        else:
            x = sigmoid(x)
    return x


def dhe_embed(ids, keys, weights, biases, bf16_points=False):
    """dh_embedder.py:205-245: hashes (NOT de-padded) -> MLP."""
    return dhe_mlp(dhe_hashes(ids, keys), weights, biases, bf16_points)


def featnet_feature_matrix(columns: Sequence[np.ndarray]) -> np.ndarray:
    """feat_dh_embedder.py:96-99 / dnn_embedder.py:63-64: every column block viewed [n, -1], L2-normalised per row,
    hstacked (no 'global' option here)."""
    n = np.asarray(columns[0]).shape[0]
    return np.hstack([l2_normalize(np.asarray(c, np.float32).reshape(n, -1)) for c in columns]).astype(np.float32)


def fdhe_embed(ids, keys, feature_mat, weights, biases, training=False, prime_pad=OOV_PRIME_PAD, bf16_points=False):
    """feat_dh_embedder.py:188-213 (`keys` given) and dnn_embedder.py:87-109 (`keys=None`): the 4-layer net on
    hstack(hashes of the ORIGINAL id, feature row of the de-padded id).  `bf16_points`: the feature inputs are rounded
    to bf16 like the hidden activations (the caller rounds the weights); hash inputs stay exact."""
    ids = np.asarray(ids, np.int64)
    feat = np.asarray(feature_mat, np.float32)[depad_ids(ids, training, prime_pad)]
    if bf16_points:
        feat = round_bf16(feat)
    x = feat if keys is None else np.hstack([dhe_hashes(ids, keys).astype(np.float32), feat])
    return dhe_mlp(x, weights, biases, bf16_points)


# ---------------------------------------------------------------------------------------
# mean / zero (mean_embedder.py:42-87, zero_embedder.py:36-60)
# ---------------------------------------------------------------------------------------
def mean_embed(table: np.ndarray, n: int) -> np.ndarray:
    """Column mean over ALL rows of the in-vocab table (incl. pad row 0), repeated n×."""
    m = np.asarray(table, np.float32).mean(axis=0, dtype=np.float32)
    return np.repeat(m.reshape(1, -1), n, axis=0).astype(np.float32)


def zero_embed(n: int, d: int) -> np.ndarray:
    return np.zeros((n, d), dtype=np.float32)


# ---------------------------------------------------------------------------------------
# Table assembly + scoring + masks + top-k
# (bpr.py:48-125,151-156; directau.py:107-198; inductive/evaluator.py:91-94; evaluator/collector.py:153-167)
# ---------------------------------------------------------------------------------------
def assemble_rows(ids: np.ndarray, n_old: int, iv_table: np.ndarray,
                  oov_embed: Callable[[np.ndarray], np.ndarray]) -> np.ndarray:
    """bpr.py:94-125: E[i] = table[id] if id < n_old else embedder(id)."""
    ids = np.asarray(ids, np.int64)
    d = iv_table.shape[1]
    out = np.zeros((ids.shape[0], d), dtype=np.float32)
    iv = ids < n_old
    out[iv] = np.asarray(iv_table, np.float32)[ids[iv]]
    if (~iv).any():
        out[~iv] = oov_embed(ids[~iv])
    return out


def full_sort_scores(user_e: np.ndarray, item_e: np.ndarray) -> np.ndarray:
    """bpr.py:155 / directau.py:197: raw dot products, no normalisation."""
    return (np.asarray(user_e, np.float32) @ np.asarray(item_e, np.float32).T).astype(np.float32)


def mask_scores(scores: np.ndarray, hist_u: Optional[np.ndarray] = None,
                hist_i: Optional[np.ndarray] = None) -> np.ndarray:
    """inductive/evaluator.py:91-94: pad item column 0 and the user's history -> -inf."""
    s = np.array(scores, dtype=np.float32, copy=True)
    s[:, 0] = -np.inf
    if hist_u is not None and len(hist_u):
        s[np.asarray(hist_u, np.int64), np.asarray(hist_i, np.int64)] = -np.inf
    return s


def segment_mask(scores: np.ndarray, n_old_items: int, keep_old: Optional[bool]) -> np.ndarray:
    """inductive/collector_filter.py:172-175: the filtered collectors blank one item
    segment (ids >= n_old when keeping old items, ids < n_old when keeping new)."""
    s = np.array(scores, dtype=np.float32, copy=True)
    if keep_old is None:
        return s
    if keep_old:
        s[:, n_old_items:] = -np.inf
    else:
        s[:, :n_old_items] = -np.inf
    return s


def order_key(scores: np.ndarray) -> np.ndarray:
    """torch.topk orders NaN above every number (an all-zero LSH multi-hot row gives a
    NaN embedding, lsh_embedder.py:158, hence NaN scores): map NaN -> +inf for ranking."""
    s = np.asarray(scores, np.float32)
    return np.where(np.isnan(s), np.float32(np.inf), s)


def topk(scores: np.ndarray, k: int):
    """evaluator/collector.py:153-159: torch.topk(scores, k).  torch leaves tie order
    unspecified (and FilteredCollector randomises it, filtered_collector.py:38-48);
    this restatement fixes the rule (score desc, index asc) — compare with
    `topk_sets_match`, which treats elements tied at the k-th score as interchangeable."""
    s = np.asarray(scores, np.float32)
    q, n = s.shape
    k = min(k, n)
    key = order_key(s)
    # lexsort: last key is primary
    order = np.lexsort((np.broadcast_to(np.arange(n), s.shape), -key), axis=1)[:, :k]
    vals = np.take_along_axis(s, order, axis=1)
    return vals, order.astype(np.int64)


def topk_sets_match(scores: np.ndarray, idx: np.ndarray, k: int, rtol: float = 0.0, atol: float = 0.0):
    """Tie-aware set equality of a candidate top-k index matrix against `scores`.

    Returns (ok, message).  A candidate row is accepted iff it has k distinct
    indices, every index whose score is strictly above the k-th best score
    (beyond tolerance) is present, and every chosen index has a score >= the k-th
    best score (within tolerance).
    """
    s = order_key(scores)
    idx = np.asarray(idx, np.int64)
    q, n = s.shape
    k = min(k, n)
    if idx.shape != (q, k):
        return False, f"shape {idx.shape} != {(q, k)}"
    part = np.sort(s, axis=1)[:, ::-1]
    kth = part[:, k - 1]
    for r in range(q):
        row = idx[r]
        if len(set(row.tolist())) != k:
            return False, f"row {r}: duplicate indices"
        if row.min() < 0 or row.max() >= n:
            return False, f"row {r}: index out of range"
        tol = atol + rtol * abs(float(kth[r])) if np.isfinite(kth[r]) else 0.0
        chosen = s[r, row]
        if np.isfinite(kth[r]) and (chosen < kth[r] - tol).any():
            return False, f"row {r}: chose a score below the k-th best"
        must = np.nonzero(s[r] > kth[r] + tol)[0]
        if not set(must.tolist()).issubset(set(row.tolist())):
            return False, f"row {r}: missed an item strictly above the k-th score"
    return True, "ok"


def collector_hits(topk_idx: np.ndarray, positive_u: np.ndarray, positive_i: np.ndarray, n_items: int):
    """evaluator/collector.py:160-166: [hits(k) | pos_len] int matrix ('rec.topk')."""
    q = topk_idx.shape[0]
    pos = np.zeros((q, n_items), dtype=np.int32)
    pos[np.asarray(positive_u, np.int64), np.asarray(positive_i, np.int64)] = 1
    pos_len = pos.sum(axis=1, keepdims=True)
    hits = np.take_along_axis(pos, np.asarray(topk_idx, np.int64), axis=1)
    return np.concatenate([hits, pos_len], axis=1)


def merge_topk(cand_scores: np.ndarray, cand_idx: np.ndarray, k: int):
    """Shard merge (no reference counterpart — SURVEY §8e): [G,Q,k] candidates ->
    global top-k with the (score desc, index asc) rule."""
    g, q, kk = cand_scores.shape
    s = np.transpose(cand_scores, (1, 0, 2)).reshape(q, g * kk)
    i = np.transpose(cand_idx, (1, 0, 2)).reshape(q, g * kk)
    order = np.lexsort((i, -order_key(s)), axis=1)[:, :k]
    return np.take_along_axis(s, order, axis=1), np.take_along_axis(i, order, axis=1)


# ---------------------------------------------------------------------------------------
# Context models: token gather + OOV overwrite
# (model/abstract_recommender.py:794-842, model/layers.py:150-153, 1634-1693)
# ---------------------------------------------------------------------------------------
def embed_token_fields(token_fields: np.ndarray, offsets: np.ndarray, table: np.ndarray,
                       n_users: int, n_items: int,
                       embed_user: Callable[[np.ndarray], np.ndarray],
                       embed_item: Callable[[np.ndarray], np.ndarray],
                       uid_idx: int = 0, iid_idx: int = 1) -> np.ndarray:
    """[B, fields] ids -> [B, fields, D]; OOV user/item ids are looked up as id 0 and the
    resulting row is then overwritten by the embedder output."""
    tf = np.array(token_fields, dtype=np.int64, copy=True)
    user_ids = tf[:, uid_idx].copy()
    item_ids = tf[:, iid_idx].copy()
    oov_u = user_ids >= n_users
    oov_i = item_ids >= n_items
    tf[oov_u, uid_idx] = 0
    tf[oov_i, iid_idx] = 0
    out = np.asarray(table, np.float32)[tf + np.asarray(offsets, np.int64).reshape(1, -1)]
    if oov_u.any():
        out[oov_u, uid_idx] = embed_user(user_ids[oov_u])
    if oov_i.any():
        out[oov_i, iid_idx] = embed_item(item_ids[oov_i])
    return out


def first_order_token_sum(token_fields, offsets, table1, n_users, n_items, embed_user, embed_item):
    """layers.py:1634-1693: same with D = 1 tables, then sum over fields -> [B, 1, 1]."""
    e = embed_token_fields(token_fields, offsets, table1, n_users, n_items, embed_user, embed_item)
    return e.sum(axis=1, keepdims=True, dtype=np.float32)


# ---------------------------------------------------------------------------------------
# inductive_mapper=random integer hashes (random_mapper.py:70-130) — "next" row f3
# ---------------------------------------------------------------------------------------
def _wrap_i64(x: np.ndarray) -> np.ndarray:
    return x.astype(np.int64)


def mapper_hash(ids: np.ndarray, n_buckets: int, fn: str) -> np.ndarray:
    """int64 tensors, wrapping multiplies, arithmetic >>, Python-sign %."""
    x = np.asarray(ids, dtype=np.int64).copy()
    with np.errstate(over="ignore"):
        if fn == "mod":
            pass
        elif fn == "fast":
            x = x ^ (x >> 16); x = x * np.int64(0x21F0AAAD)
            x = x ^ (x >> 15); x = x * np.int64(0xD35A2D97)
            x = x ^ (x >> 15)
        elif fn == "3round":
            x = x ^ (x >> 17); x = x * np.int64(0xED5AD4BB)
            x = x ^ (x >> 11); x = x * np.int64(0xAC4C1B51)
            x = x ^ (x >> 15); x = x * np.int64(0x31848BAB)
            x = x ^ (x >> 14)
        elif fn == "64bit":
            u = x.astype(np.uint64)
            u = (u ^ (u >> np.uint64(30))) * np.uint64(0xB9E5E41C6D4758BF)
            u = (u ^ (u >> np.uint64(27))) * np.uint64(0xEB113113BB49D094)
            u = u ^ (u >> np.uint64(31))
            return (u % np.uint64(n_buckets)).astype(np.int64)
        else:
            raise ValueError(f"Unknown hash function {fn}")
    return np.mod(x, np.int64(n_buckets))


def map_ids(ids: np.ndarray, n_old: int, n_buckets: int, fn: str) -> np.ndarray:
    """random_mapper.py:113-130: id if id < n_old else n_old + hash(id - n_old) % B."""
    ids = np.asarray(ids, np.int64)
    out = ids.copy()
    oov = ids >= n_old
    out[oov] = mapper_hash(ids[oov] - n_old, n_buckets, fn) + n_old
    return out


# ---------------------------------------------------------------------------------------
# DCN-V2 dense tower in eval mode
# ---------------------------------------------------------------------------------------
def dcnv2_cross(x0: np.ndarray, cross_w: Sequence[np.ndarray], cross_b: Sequence[np.ndarray], bf16_points: bool = False,
                bf16_t: bool = True) -> np.ndarray:
    """model/context_aware_recommender/dcnv2.py:120-144: x_{l+1} = x_0 * (W_l x_l + b_l) + x_l, rows of x0 [B, in].
    bf16_points: round every x_l to bf16 — and, with bf16_t, the linear's output t too (the rounding points of the
    tensor-core path: a bf16 linear followed by the elementwise tail)."""
    r = round_bf16 if bf16_points else (lambda a: a)
    rt = r if bf16_t else (lambda a: a)
    x0 = np.asarray(x0, np.float32)
    xl = x0
    for w, b in zip(cross_w, cross_b):
        t = rt((xl @ np.asarray(w, np.float32).T + np.asarray(b, np.float32).reshape(1, -1)).astype(np.float32))
        xl = r((x0 * t + xl).astype(np.float32))
    return xl


def mlp_bn_relu(x: np.ndarray, layers: Sequence[dict], bf16_points: bool = False) -> np.ndarray:
    """model/layers.py:33-92 MLPLayers(bn=True, activation='relu') in eval mode: Dropout is the identity, BatchNorm1d uses
    its running statistics: y = relu((x W^T + b - mean) / sqrt(var + eps) * gamma + beta).
    layers: dicts with w [out, in], b, bn_mean, bn_var, bn_gamma, bn_beta, bn_eps."""
    r = round_bf16 if bf16_points else (lambda a: a)
    h = np.asarray(x, np.float32)
    for L in layers:
        z = h @ np.asarray(L["w"], np.float32).T + np.asarray(L["b"], np.float32)
        z = (z - L["bn_mean"]) / np.sqrt(L["bn_var"] + L["bn_eps"]) * L["bn_gamma"] + L["bn_beta"]
        h = r(np.maximum(z, 0).astype(np.float32))
    return h


def fold_bn(L: dict):
    """Eval-mode BatchNorm folded into the preceding Linear: W' = W * (gamma / sigma), b' = (b - mean) * gamma / sigma + beta."""
    s = (L["bn_gamma"] / np.sqrt(L["bn_var"] + L["bn_eps"])).astype(np.float32)
    return (np.asarray(L["w"], np.float32) * s[:, None]).astype(np.float32), ((L["b"] - L["bn_mean"]) * s + L["bn_beta"]).astype(np.float32)


def dcnv2_forward(x0, cross_w, cross_b, mlp_layers, pred_w, pred_b, structure: str = "stacked", bf16_points: bool = False,
                  bf16_t: bool = True):
    """dcnv2.py:214-250 (mixed = False): stacked: sigmoid(predict(mlp(cross(x0)))); parallel: sigmoid(predict([cross(x0) | mlp(x0)]))."""
    c = dcnv2_cross(x0, cross_w, cross_b, bf16_points, bf16_t)
    if structure == "stacked":
        top = mlp_bn_relu(c, mlp_layers, bf16_points)
    else:
        top = np.concatenate([c, mlp_bn_relu(x0, mlp_layers, bf16_points)], axis=1)
    z = top @ np.asarray(pred_w, np.float32).reshape(-1) + np.float32(np.asarray(pred_b).reshape(-1)[0])
    return sigmoid(z.astype(np.float32))


# ------------------------------------------------------------------------------------ sampled-negative evaluation
def pair_scores(user_e: np.ndarray, item_e: np.ndarray, normalize: bool = False) -> np.ndarray:
    """model/general_recommender/bpr.py:146-149 predict: torch.mul(user_e, item_e).sum(dim=1) for aligned [P, D] rows;
    normalize: directau.py:75-78,174-181 — F.normalize(x, dim=-1) = x / max(|x|_2, 1e-12) on both sides first."""
    u, v = np.asarray(user_e, np.float32), np.asarray(item_e, np.float32)
    if normalize:
        u = u / np.maximum(np.sqrt((u * u).sum(axis=1, keepdims=True, dtype=np.float32)), np.float32(1e-12))
        v = v / np.maximum(np.sqrt((v * v).sum(axis=1, keepdims=True, dtype=np.float32)), np.float32(1e-12))
    return (u * v).sum(axis=1, dtype=np.float32)


def neg_sample_scores(origin_scores: np.ndarray, row_idx: np.ndarray, col_idx: np.ndarray, n_rows: int, n_items: int) -> np.ndarray:
    """trainer/trainer.py:559-564 = inductive/evaluator.py:127-133: scores = full((batch_user_num, tot_item_num), -inf);
    scores[row_idx, col_idx] = origin_scores (duplicate pairs carry the same score)."""
    s = np.full((n_rows, n_items), -np.inf, dtype=np.float32)
    s[np.asarray(row_idx), np.asarray(col_idx)] = np.asarray(origin_scores, np.float32)
    return s


def widedeep_forward(emb, fm, mlp_w, mlp_b, pred_w, pred_b, bf16_points: bool = False):
    """model/context_aware_recommender/widedeep.py:70-81: logits = first_order_linear + deep_predict_layer(mlp_layers(
    emb.view(B, -1))) with MLPLayers(bn=False, activation='relu') in eval mode (layers.py:33-92); predict = sigmoid(logits).
    bf16_points: round the input and every hidden activation to bf16 (the rounding points of the tensor-core path)."""
    r = round_bf16 if bf16_points else (lambda a: a)
    h = r(np.asarray(emb, np.float32).reshape(np.asarray(emb).shape[0], -1))
    for w, b in zip(mlp_w, mlp_b):
        h = r(np.maximum(h @ np.asarray(w, np.float32).T + np.asarray(b, np.float32), 0).astype(np.float32))
    deep = h @ np.asarray(pred_w, np.float32).reshape(-1) + np.float32(np.asarray(pred_b).reshape(-1)[0])
    return (np.asarray(fm, np.float32).reshape(-1) + deep).astype(np.float32)


def xdeepfm_cin(emb, conv_w, conv_b, direct: bool = False, bf16_points: bool = False):
    """model/context_aware_recommender/xdeepfm.py:134-190 (activation ReLU): per layer z = einsum("bhd,bmd->bhmd",
    X^{k-1}, X^0) viewed [B, H*M, D], kernel-size-1 Conv1d (conv_w[k]: [O, H*M]), ReLU; direct = False splits the
    output channels into (next_hidden, direct_connect) halves except for the last layer; the direct-connect parts are
    concatenated and sum-pooled over D -> [B, final_len].
    bf16_points: z and every layer output are rounded to bf16 (the tensor-core path's operand / activation precision)."""
    r = round_bf16 if bf16_points else (lambda a: a)
    x0 = np.asarray(emb, np.float32)
    B, M, D = x0.shape
    hidden, final = x0, []
    n_layers = len(conv_w)
    for i, (w, b) in enumerate(zip(conv_w, conv_b)):
        z = r((hidden[:, :, None, :] * x0[:, None, :, :]).reshape(B, -1, D).astype(np.float32))
        w = np.asarray(w, np.float32)
        out = np.einsum("oc,bcd->bod", w.astype(np.float64), z.astype(np.float64)) + np.asarray(b, np.float64)[None, :, None]
        out = r(np.maximum(out, 0).astype(np.float32))
        size = w.shape[0]
        if direct:
            direct_connect, hidden = out, out
        elif i != n_layers - 1:
            hidden, direct_connect = out[:, : size // 2], out[:, size // 2:]
        else:
            direct_connect = out
        final.append(direct_connect)
    return np.concatenate(final, axis=1).astype(np.float64).sum(axis=-1).astype(np.float32)


def xdeepfm_forward(emb, fm, conv_w, conv_b, lin_w, lin_b, mlp_w, mlp_b, direct: bool = False, bf16_points: bool = False):
    """xdeepfm.py:192-207: logits = first_order_linear + cin_linear(CIN(emb)) + mlp_layers(emb.view(B, -1)), the MLP being
    MLPLayers(sizes + [1]) with ReLU after EVERY Linear, the 1-wide last one included (layers.py:60-75)."""
    r = round_bf16 if bf16_points else (lambda a: a)
    emb = np.asarray(emb, np.float32)
    cin = xdeepfm_cin(emb, conv_w, conv_b, direct, bf16_points)
    cin = cin.astype(np.float64) @ np.asarray(lin_w, np.float64).reshape(-1) + float(np.asarray(lin_b).reshape(-1)[0])
    h = r(emb.reshape(emb.shape[0], -1))
    for l, (w, b) in enumerate(zip(mlp_w, mlp_b)):
        h = np.maximum(h.astype(np.float64) @ np.asarray(w, np.float64).T + np.asarray(b, np.float64), 0).astype(np.float32)
        if l != len(mlp_w) - 1:
            h = r(h)
    return (np.asarray(fm, np.float32).reshape(-1) + cin.astype(np.float32) + h[:, 0]).astype(np.float32)


# ------------------------------------------------------------------------------------ training-mode backward
def lsh_embed_backward(multihot: np.ndarray, g: np.ndarray) -> np.ndarray:
    """d/dW of new_embed = (H @ W) / H.sum(1) (lsh_embedder.py:156-158) given g = d loss / d new_embed: H^T (g / |H|),
    which is what autograd computes (an all-zero row gives 0 * inf = NaN on every bucket)."""
    H = np.asarray(multihot, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (H.T.astype(np.float64) @ (np.asarray(g, np.float64) / H.sum(axis=1, keepdims=True).astype(np.float64))).astype(np.float32)


def scatter_add_rows(g: np.ndarray, idx: np.ndarray, rows: int) -> np.ndarray:
    """nn.Embedding backward: d table[idx[i]] += g[i] for 0 <= idx[i] < rows."""
    out = np.zeros((rows, np.asarray(g).shape[1]), np.float64)
    idx = np.asarray(idx, np.int64)
    ok = (idx >= 0) & (idx < rows)
    np.add.at(out, idx[ok], np.asarray(g, np.float64)[ok])
    return out.astype(np.float32)
