"""Shared helpers for the parity tests: build the oracle's outputs for a seeded case,
and compare bit matrices / embeddings / top-k sets with the tolerances the north
star states (written out here, once):

* hash bits, bucket ids, DHE hash ids, top-k index SETS: exact, except LSH bits whose
  reference projection has |x| < TIE_EPS (counted and reported, never silently dropped);
* fp32 embeddings and scores: rtol 1e-5 (+ atol 1e-6 for values that cancel to ~0);
* bf16 path: rtol 1e-3 (+ atol 1e-3·scale) against the oracle evaluated at the same
  bf16 rounding points.
"""
from __future__ import annotations

import os

import numpy as np

import cases
from oracle import oracle as o

TIE_EPS = 1e-6          # north_star: "projection-sign ties with |x|<1e-6 are counted and reported"
FP32_RTOL, FP32_ATOL = 1e-5, 1e-6
BF16_RTOL = 1e-3

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)


def unpack_bits(packed: np.ndarray, n_bits: int) -> np.ndarray:
    return np.unpackbits(packed, axis=1, bitorder="little")[:, :n_bits]


def words_to_bits(words: np.ndarray, n_bits: int) -> np.ndarray:
    """uint32 [n, ceil(B/32)] (bit b of word w = plane 32*w+b) -> uint8 [n, B]."""
    w = np.ascontiguousarray(words).view(np.uint8)
    return np.unpackbits(w, axis=1, bitorder="little")[:, :n_bits]


def tie_positions(near_rows, near_cols, near_vals, eps=TIE_EPS):
    m = np.abs(near_vals) < eps
    return set(zip(near_rows[m].tolist(), near_cols[m].tolist()))


def check_bits(got: np.ndarray, want: np.ndarray, ties: set) -> int:
    """Bit-exact modulo reported ties.  Returns the number of tie positions that differ."""
    assert got.shape == want.shape, (got.shape, want.shape)
    rr, cc = np.nonzero(got != want)
    bad = [(r, c) for r, c in zip(rr.tolist(), cc.tolist()) if (r, c) not in ties]
    assert not bad, f"{len(bad)} bit mismatches outside the |x|<{TIE_EPS} tie class, first {bad[:5]}"
    return len(rr)


def rows_with_bit_diffs(got: np.ndarray, want: np.ndarray) -> np.ndarray:
    return np.nonzero((got != want).any(axis=1))[0]


def assert_close(got, want, rtol=FP32_RTOL, atol=FP32_ATOL, skip_rows=None, what=""):
    got = np.asarray(got, np.float32)
    want = np.asarray(want, np.float32)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if skip_rows is not None and len(skip_rows):
        keep = np.ones(got.shape[0], bool)
        keep[np.asarray(skip_rows)] = False
        got, want = got[keep], want[keep]
    both_nan = np.isnan(got) & np.isnan(want)
    same_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    with np.errstate(invalid="ignore"):
        ok = both_nan | same_inf | (np.abs(got - want) <= atol + rtol * np.abs(want))
    if not ok.all():
        bad = np.argwhere(~ok)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} / {ok.size} elements out of tolerance "
                             f"(rtol={rtol}, atol={atol}); first at {i}: got {got[i]!r} want {want[i]!r}")


# ---------------------------------------------------------------------------------------
def oracle_retrieval(case: cases.RetrievalCase, inp: dict) -> dict:
    """Everything the reference produces for a retrieval case, from the oracle."""
    out = {}
    ufm = o.feature_matrix(inp["user_cols"], case.normalization)
    ifm = o.feature_matrix(inp["item_cols"], case.normalization)
    out["user_feature_mat"], out["item_feature_mat"] = ufm, ifm
    oov_users = np.arange(case.n_old_users, case.n_all_users)
    oov_items = np.arange(case.n_old_items, case.n_all_items)

    if case.embedder == "lsh":
        eu = lambda ids: o.lsh_embed(ufm, ids, inp["user_planes"], inp["user_oov"])
        ei = lambda ids: o.lsh_embed(ifm, ids, inp["item_planes"], inp["item_oov"])
        out["user_bits"] = o.lsh_multihot(ufm, oov_users, inp["user_planes"]).astype(np.uint8)
        out["item_bits"] = o.lsh_multihot(ifm, oov_items, inp["item_planes"]).astype(np.uint8)
    elif case.embedder == "slsh":
        eu = lambda ids: o.slsh_embed(ufm, ids, inp["user_planes"], inp["user_oov"])
        ei = lambda ids: o.slsh_embed(ifm, ids, inp["item_planes"], inp["item_oov"])
        out["user_bucket_ids"] = o.slsh_ids(ufm, oov_users, inp["user_planes"], case.B_user)
        out["item_bucket_ids"] = o.slsh_ids(ifm, oov_items, inp["item_planes"], case.B_item)
    elif case.embedder == "mean":
        eu = lambda ids: o.mean_embed(inp["user_table"], len(ids))
        ei = lambda ids: o.mean_embed(inp["item_table"], len(ids))
    elif case.embedder == "zero":
        eu = lambda ids: o.zero_embed(len(ids), case.D)
        ei = lambda ids: o.zero_embed(len(ids), case.D)
    else:
        raise ValueError(case.embedder)
    out["oov_user_emb"] = eu(oov_users)
    out["oov_item_emb"] = ei(oov_items)
    out["user_e"] = o.assemble_rows(inp["users"], case.n_old_users, inp["user_table"], eu)
    out["all_item_e"] = o.assemble_rows(np.arange(case.n_all_items), case.n_old_items, inp["item_table"], ei)
    out["all_user_e"] = o.assemble_rows(np.arange(case.n_all_users), case.n_old_users, inp["user_table"], eu)
    out["scores_raw"] = o.full_sort_scores(out["user_e"], out["all_item_e"])
    out["scores_masked"] = o.mask_scores(out["scores_raw"], inp["hist_u"], inp["hist_i"])
    out["topk_vals"], out["topk_idx"] = o.topk(out["scores_masked"], case.k)
    return out


def history_csr(hist_u: np.ndarray, hist_i: np.ndarray, q: int):
    """(row, item) history pairs (general_dataloader.py:270-292) -> CSR rowptr/cols."""
    order = np.lexsort((hist_i, hist_u))
    hu, hi = hist_u[order], hist_i[order]
    rowptr = np.zeros(q + 1, dtype=np.int64)
    np.add.at(rowptr, hu + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), hi.astype(np.int32)


def pairs_to_csr_host(rows_idx, cols_idx, Q, col_ranges=None):
    """Host restatement (torch index ops, any device) of the contract of oov_pairs_to_csr / ops.pairs_to_csr — test
    infrastructure: int32 rowptr [Q + 1], int32 cols ascending per row; rows outside [0, Q) are padding and dropped;
    col_ranges = ((lo0, hi0), (lo1, hi1)) keeps the items of the two ranges and rewrites them to local shard rows."""
    import torch
    rows_idx = rows_idx.to(torch.int64)
    cols_idx = cols_idx.to(torch.int64)
    if rows_idx.numel() == 0:
        return torch.zeros(Q + 1, dtype=torch.int32, device=rows_idx.device), torch.zeros(0, dtype=torch.int32, device=rows_idx.device)
    drop = (rows_idx < 0) | (rows_idx >= Q)
    if col_ranges is not None:
        (a0, b0), (a1, b1) = col_ranges
        in0 = (cols_idx >= a0) & (cols_idx < b0)
        in1 = (cols_idx >= a1) & (cols_idx < b1)
        cols_idx = torch.where(in0, cols_idx - a0, cols_idx - a1 + (b0 - a0))
        drop = drop | ~(in0 | in1)
    rows_idx = torch.where(drop, torch.full_like(rows_idx, Q), rows_idx)
    cols_idx = torch.where(drop, torch.zeros_like(cols_idx), cols_idx)
    key = rows_idx * (1 << 32) + cols_idx
    key, _ = torch.sort(key)
    r = key >> 32
    c = (key & 0xFFFFFFFF).to(torch.int32)
    rowptr = torch.searchsorted(r, torch.arange(Q + 1, device=key.device, dtype=torch.int64))
    return rowptr.to(torch.int32), c


def sampled_rows_match(dense_seg: np.ndarray, got_s: np.ndarray, got_i: np.ndarray, k: int, rtol=1e-5, atol=1e-6):
    """Per row of a sampled-candidate top-k: the first n = min(k, #finite candidates) slots are a tie-aware top-n of the
    row's dense (-inf elsewhere) scores with matching values, the remaining slots are (-inf, -1)."""
    for r in range(dense_seg.shape[0]):
        fin = np.isfinite(dense_seg[r]) | np.isnan(dense_seg[r])
        n = min(k, int(fin.sum()))
        assert (got_i[r, n:] == -1).all() and np.isneginf(got_s[r, n:]).all(), (r, n, got_i[r])
        if n == 0:
            continue
        ok, msg = o.topk_sets_match(dense_seg[r:r + 1], got_i[r:r + 1, :n], n, rtol=rtol, atol=atol)
        assert ok, f"row {r}: {msg}"
        np.testing.assert_allclose(got_s[r, :n], dense_seg[r, got_i[r, :n]], rtol=rtol, atol=atol)
        key = o.order_key(got_s[r, :n])
        assert (key[:-1] >= key[1:]).all()
        tie = key[:-1] == key[1:]
        assert (got_i[r, :n][:-1][tie] < got_i[r, :n][1:][tie]).all()        # (score desc, id asc)
