"""GPU parity (B200): LSH / SLSH / mean / zero embedders, fused assemble, dense scores, fused
score+mask+top-k and the collectors — product path (through the C-ABI) vs the golden fixtures of the
unmodified reference and vs the oracle on the same seeded inputs."""
import os

import numpy as np
import pytest
import torch

import cases
import parity_util as pu
from oracle import oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


@pytest.mark.parametrize("name", list(cases.CASES))
def test_retrieval_fp32(name, G):
    from oov_b200 import ops
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    g = pu.load_golden(name)
    ora = pu.oracle_retrieval(case, inp)
    cfg, emb, model = G.build_retrieval(case, inp)
    launches0 = ops.launch_count()

    # checkpoint key names are part of the drop-in contract
    assert set(g["state_dict_keys"].tolist()) <= set(model.state_dict().keys())

    oov_users = torch.arange(case.n_old_users, case.n_all_users, device=G.DEV)
    oov_items = torch.arange(case.n_old_items, case.n_all_items, device=G.DEV)
    skip_u = skip_i = np.zeros(0, np.int64)
    n_tie_diffs = 0
    if case.embedder in ("lsh", "slsh"):
        pu.assert_close(emb.user_feature_mat.cpu().numpy(), g["user_feature_mat"], rtol=2e-6, atol=1e-7, what="user_feature_mat")
        pu.assert_close(emb.item_feature_mat.cpu().numpy(), g["item_feature_mat"], rtol=2e-6, atol=1e-7, what="item_feature_mat")
    if case.embedder == "lsh":
        for side, ids, lsh, fm, B in (("user", oov_users, emb.user_lsh, emb.user_feature_mat, case.B_user),
                                      ("item", oov_items, emb.item_lsh, emb.item_feature_mat, case.B_item)):
            words = emb._hash_node_packed(ids, lsh, fm).cpu().numpy().view(np.uint32)
            got = pu.words_to_bits(words, B)
            want = pu.unpack_bits(g[f"{side}_bits"], B)
            ties = pu.tie_positions(g[f"{side}_near_rows"], g[f"{side}_near_cols"], g[f"{side}_near_vals"])
            n_tie_diffs += pu.check_bits(got, want, ties)          # bit-exact modulo the reported |x|<1e-6 class
            if side == "user":
                skip_u = pu.rows_with_bit_diffs(got, want)
            else:
                skip_i = pu.rows_with_bit_diffs(got, want)
            # the reference-shaped {0,1} fp32 matrix
            dense = (emb._hash_users(ids) if side == "user" else emb._hash_items(ids)).cpu().numpy()
            assert dense.dtype == np.float32 and (dense.astype(np.uint8) == got).all()
        # ties are counted on the device and reported
        n_reported = int(emb.tie_count.item())
        n_expected = sum(len(pu.tie_positions(g[f"{s}_near_rows"], g[f"{s}_near_cols"], g[f"{s}_near_vals"])) for s in ("user", "item"))
        print(f"[{name}] projection-sign ties: kernel reported {n_reported} (x2 calls), reference has {n_expected}, bit diffs {n_tie_diffs}")
        assert n_reported >= n_tie_diffs
    if case.embedder == "slsh":
        for side, ids in (("user", oov_users), ("item", oov_items)):
            got = (emb._hash_users(ids) if side == "user" else emb._hash_items(ids)).cpu().numpy()
            want = g[f"{side}_bucket_ids"]
            assert got.dtype == np.int64
            diff = np.nonzero(got != want)[0]
            tie_rows = set(g[f"{side}_near_rows"][np.abs(g[f"{side}_near_vals"]) < pu.TIE_EPS].tolist())
            assert set(diff.tolist()) <= tie_rows, f"{side}: bucket ids differ outside the tie class at rows {diff[:5]}"
            if side == "user":
                skip_u = diff
            else:
                skip_i = diff

    # plugin API: embed_user_ids / embed_item_ids
    eu = emb.embed_user_ids(oov_users.clone(), model).cpu().numpy()
    ei = emb.embed_item_ids(oov_items.clone(), model).cpu().numpy()
    pu.assert_close(eu, g["oov_user_emb"], skip_rows=skip_u, what="oov_user_emb vs golden")
    pu.assert_close(ei, g["oov_item_emb"], skip_rows=skip_i, what="oov_item_emb vs golden")
    pu.assert_close(eu, ora["oov_user_emb"], skip_rows=skip_u, what="oov_user_emb vs oracle")

    # model API: fused in-vocab gather + OOV embed
    users = G.t(inp["users"])
    user_e = model.get_user_embedding(users.clone()).cpu().numpy()
    item_range = torch.arange(case.n_all_items, device=G.DEV)
    all_item_e = model.get_item_embedding(item_range).cpu().numpy()
    if len(skip_u) == 0 and len(skip_i) == 0:
        pu.assert_close(user_e, g["user_e"], what="user_e")
        pu.assert_close(all_item_e, g["all_item_e"], what="all_item_e")
        # in-vocab rows are copies: bit-exact
        assert (all_item_e[: case.n_old_items] == inp["item_table"]).all()

        fin = np.isfinite(g["scores_raw"])
        scale = float(np.abs(g["scores_raw"][fin]).max()) if fin.any() else 1.0
        inter = {"user_id": users.clone()}
        dense = model.ind_full_sort_predict(inter, item_range).view(-1, case.n_all_items).cpu().numpy()
        pu.assert_close(dense, g["scores_raw"], rtol=1e-5, atol=1e-5 * scale, what="ind_full_sort_predict")

        # fused score + pad/history mask + top-k
        hist = (G.t(inp["hist_u"]), G.t(inp["hist_i"]))
        s, idx = model.full_sort_topk(inter, case.k, n_total_items=case.n_all_items, history_index=hist)
        s, idx = s.cpu().numpy(), idx.cpu().numpy()
        ok, msg = o.topk_sets_match(ora["scores_masked"], idx, case.k, rtol=1e-5, atol=1e-6 * scale)
        assert ok, "fused top-k vs oracle: " + msg
        gold_masked = o.mask_scores(g["scores_raw"], inp["hist_u"], inp["hist_i"])
        ok, msg = o.topk_sets_match(gold_masked, idx, case.k, rtol=1e-5, atol=1e-6 * scale)
        assert ok, "fused top-k vs reference scores: " + msg
        # returned scores are the scores of the returned ids, ordered (score desc, id asc)
        picked = np.take_along_axis(ora["scores_masked"], idx, axis=1)
        pu.assert_close(s, picked, rtol=1e-5, atol=1e-5 * scale, what="top-k scores")
        key = o.order_key(s)
        assert (key[:, :-1] >= key[:, 1:]).all()
        tie = key[:, :-1] == key[:, 1:]
        assert (idx[:, :-1][tie] < idx[:, 1:][tie]).all()

        # dense top-k entry point on the reference-shaped matrix
        masked_dev = ops.fullsort_scores(G.t(ora["user_e"]), G.t(ora["all_item_e"]), mask_pad=True,
                                         hist=ops.pairs_to_csr(hist[0], hist[1], case.Q))
        pu.assert_close(masked_dev.cpu().numpy(), ora["scores_masked"], rtol=1e-5, atol=1e-5 * scale, what="masked dense scores")
        ds, di = ops.dense_topk(masked_dev, case.k)
        ok, msg = o.topk_sets_match(masked_dev.cpu().numpy(), di.cpu().numpy(), case.k)
        assert ok, "dense_topk: " + msg

        # collector 'rec.topk' = [hits | pos_len]
        srt = -np.sort(-o.order_key(ora["scores_masked"]), axis=1)
        with np.errstate(invalid="ignore"):
            clear = (srt[:, case.k - 1] - srt[:, case.k]) > 1e-5 * scale
        rp, pc = ops.pairs_to_csr(G.t(inp["pos_u"]), G.t(inp["pos_i"]), case.Q)
        hits = ops.topk_hits(torch.from_numpy(idx).to(G.DEV), rp, pc).cpu().numpy()
        assert (hits[:, -1] == g["collector_overall"][:, -1]).all()
        assert (hits[clear].sum(1) == g["collector_overall"][clear].sum(1)).all()
        want_hits = o.collector_hits(idx, inp["pos_u"], inp["pos_i"], case.n_all_items)
        assert (hits == want_hits).all()
    assert ops.launch_count() > launches0


@pytest.mark.parametrize("name", ["bpr_lsh_ml100k", "directau_slsh", "bpr_mean"])
def test_inductive_evaluator_collectors(name, G):
    """The 7 collectors of inductive/evaluator.py:39-56 from ONE scoring pass (old-items + new-items launches, merged for
    the all-items list) and one hits kernel, vs the reference's FilteredCollector outputs (each fed un-aliased scores),
    in reference_compat mode; eval_batch runs under torch's sync-debug mode: no host synchronisation."""
    import oov_b200
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    g = pu.load_golden(name)
    ora = pu.oracle_retrieval(case, inp)
    cfg, emb, model = G.build_retrieval(case, inp)
    ev = oov_b200.InductiveEvaluator(model, cfg, case.n_old_users, case.n_old_items, reference_compat=True)
    ev.tot_item_num = case.n_all_items
    batch = ({"user_id": G.t(inp["users"])}, (G.t(inp["hist_u"]), G.t(inp["hist_i"])), G.t(inp["pos_u"]), G.t(inp["pos_i"]))
    table = model.build_item_table(case.n_all_items)
    ev.eval_batch(batch, item_table=table)                     # warm-up: workspaces, lazy initialisation
    ev = oov_b200.InductiveEvaluator(model, cfg, case.n_old_users, case.n_old_items, reference_compat=True)
    ev.tot_item_num = case.n_all_items
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")                    # eval_batch must not synchronise with the host
    try:
        res = ev.eval_batch(batch, item_table=table)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    scale = float(np.abs(ora["scores_raw"][np.isfinite(ora["scores_raw"])]).max())
    for cname in oov_b200.evaluator.COLLECTORS:
        key = f"collector_{cname}"
        if key not in g.files:
            assert cname not in res
            continue
        got, want = res[cname].cpu().numpy(), g[key]
        assert got.shape == want.shape, (cname, got.shape, want.shape)
        assert (got[:, -1] == want[:, -1]).all(), cname          # pos_len column
        # rows whose k-th / (k+1)-th scores are clearly separated inside the collector's item segment
        ru, ri = oov_b200.evaluator.COLLECTORS[cname]
        # collector_filter.py:172-175 selects the blanked segment from return_old_USERS (sic)
        seg = o.segment_mask(ora["scores_masked"], case.n_old_items, None if ri is None else ru)
        rows = g[f"{key}_rows"] if f"{key}_rows" in g.files else np.arange(case.Q)
        srt = -np.sort(-o.order_key(seg[rows]), axis=1)
        with np.errstate(invalid="ignore"):
            clear = (srt[:, case.k - 1] - srt[:, case.k]) > 1e-5 * scale
        assert (got[clear].sum(1) == want[clear].sum(1)).all(), cname
        assert (got[clear] == want[clear]).all(), cname


@pytest.mark.parametrize("name", ["bpr_lsh_ml100k", "directau_slsh"])
def test_retrieval_bf16_tables(name, G):
    """bf16 tables (config 2/5 precision): kernel vs the oracle evaluated on the SAME bf16-rounded
    tables, rtol 1e-3; top-k sets exact w.r.t. those scores up to that tolerance."""
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    ora = pu.oracle_retrieval(case, inp)
    cfg, emb, model = G.build_retrieval(case, inp, table_dtype="bfloat16")
    users = G.t(inp["users"])
    item_tab = model.build_item_table(case.n_all_items)
    assert item_tab.dtype == torch.bfloat16
    # the bf16 table is the rounding of the fp32 embedding (skip rows touched by sign ties)
    want_tab = o.round_bf16(ora["all_item_e"])
    got_tab = item_tab.float().cpu().numpy()
    close = np.isclose(got_tab, want_tab, rtol=2 ** -7, atol=1e-6, equal_nan=True).all(axis=1)
    assert close.mean() > 0.999
    user_e = model._assemble("user", users, out_dtype=torch.bfloat16).float().cpu().numpy()
    scores = o.mask_scores(o.full_sort_scores(user_e, got_tab), inp["hist_u"], inp["hist_i"])
    hist = (G.t(inp["hist_u"]), G.t(inp["hist_i"]))
    s, idx = model.full_sort_topk({"user_id": users}, case.k, n_total_items=case.n_all_items, history_index=hist)
    scale = float(np.abs(scores[np.isfinite(scores)]).max())
    ok, msg = o.topk_sets_match(scores, idx.cpu().numpy(), case.k, rtol=pu.BF16_RTOL, atol=1e-3 * scale * 1e-2)
    assert ok, msg
    picked = np.take_along_axis(scores, idx.cpu().numpy(), axis=1)
    pu.assert_close(s.cpu().numpy(), picked, rtol=pu.BF16_RTOL, atol=1e-5 * scale, what="bf16 top-k scores")


def test_lsh_training_mode_and_strided_io(G):
    """prime-pad de-padding (lsh_embedder.py:153-155, mutates the caller's ids) and strided ids / out."""
    case = cases.CASES["bpr_lsh_ml100k"]
    inp = cases.retrieval_inputs(case)
    g = pu.load_golden(case.name)
    cfg, emb, model = G.build_retrieval(case, inp)
    emb.set_train()
    padded = torch.arange(case.n_old_items, case.n_old_items + 16, device=G.DEV) + cases.OOV_PRIME_PAD
    got = emb.embed_item_ids(padded, model).cpu().numpy()
    pu.assert_close(got, g["oov_item_emb_train16"], what="train-mode emb")
    assert (padded.cpu().numpy() == g["padded_after"]).all()
    # model-level call in training mode: padded ids are OOV (>= n_items) and hash the ORIGINAL row
    padded2 = torch.arange(case.n_old_items, case.n_old_items + 16, device=G.DEV) + cases.OOV_PRIME_PAD
    got2 = model.get_item_embedding(padded2).cpu().numpy()
    pu.assert_close(got2, g["oov_item_emb_train16"], what="train-mode emb via model")
    emb.set_eval()

    # strided ids (column of a [n, 3] matrix) and strided out (column 1 of [n, 2, D])
    ids2d = torch.zeros((40, 3), dtype=torch.int64, device=G.DEV)
    ids2d[:, 2] = torch.arange(case.n_old_items - 20, case.n_old_items + 20, device=G.DEV)
    out3 = torch.full((40, 2, case.D), 7.0, device=G.DEV)
    emb.assemble_rows("item", ids2d[:, 2], model, case.n_old_items, None, out=out3[:, 1, :])
    out3 = out3.cpu().numpy()
    assert (out3[:, 0, :] == 7.0).all() and (out3[:20, 1, :] == 7.0).all()      # in-vocab rows untouched (iv_table=None)
    pu.assert_close(out3[20:, 1, :], g["oov_item_emb"][:20], what="strided out")


def test_empty_and_error_behaviour(G):
    from oov_b200 import ops
    case = cases.CASES["bpr_slsh_odd"]
    inp = cases.retrieval_inputs(case)
    cfg, emb, model = G.build_retrieval(case, inp)
    empty = torch.zeros(0, dtype=torch.int64, device=G.DEV)
    assert emb.embed_item_ids(empty, model).shape == (0, case.D)
    assert model.get_item_embedding(empty).shape == (0, case.D)
    with pytest.raises(RuntimeError):                 # CPU tensors: no fallback
        emb.embed_item_ids(torch.arange(3), model)
    with pytest.raises(ValueError):                   # wrong id dtype
        emb.embed_item_ids(torch.arange(3, device=G.DEV, dtype=torch.int32), model)
    users = torch.randn(4, 24, device=G.DEV)
    with pytest.raises(ValueError):                   # k out of range
        ops.fullsort_topk(users, torch.randn(10, 24, device=G.DEV), 500)
    # fewer items than k: tail is (-inf, -1)
    s, i = ops.fullsort_topk(torch.randn(4, 32, device=G.DEV), torch.randn(3, 32, device=G.DEV), 5, mask_pad=False)
    assert (i[:, 3:] == -1).all() and torch.isinf(s[:, 3:]).all() and (i[:, :3] >= 0).all()


@pytest.mark.parametrize("name", ["bpr_lsh_ml100k", "directau_slsh"])
def test_inductive_evaluator_default_mode_vs_oracle(name, G):
    """reference_compat=False (the default): a collector with an item filter ranks ITS item segment and hits are compared
    in global ids.  Expected rows from the oracle's masked scores: segment mask by return_old_items, top-k, hit flags
    against the item-filtered positives, rows = users passing the user filter that own such a positive.  Also checks that
    the accumulated collectors return the same rows (one host copy at read time)."""
    import oov_b200
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    ora = pu.oracle_retrieval(case, inp)
    cfg, emb, model = G.build_retrieval(case, inp)
    ev = oov_b200.InductiveEvaluator(model, cfg, case.n_old_users, case.n_old_items)
    assert ev.reference_compat is False
    ev.tot_item_num = case.n_all_items
    batch = ({"user_id": G.t(inp["users"])}, (G.t(inp["hist_u"]), G.t(inp["hist_i"])), G.t(inp["pos_u"]), G.t(inp["pos_i"]))
    res = ev.eval_batch(batch)
    scale = float(np.abs(ora["scores_raw"][np.isfinite(ora["scores_raw"])]).max())
    users, pos_u, pos_i = inp["users"], inp["pos_u"], inp["pos_i"]
    old_user = users < case.n_old_users
    for cname, (ru, ri) in oov_b200.evaluator.COLLECTORS.items():
        seg = o.segment_mask(ora["scores_masked"], case.n_old_items, ri)
        pm = np.ones_like(pos_u, dtype=bool)
        if ri is not None:
            pm &= (pos_i < case.n_old_items) if ri else (pos_i >= case.n_old_items)
        urow = np.ones(case.Q, dtype=bool) if ru is None else (old_user if ru else ~old_user)
        pos_len = np.bincount(pos_u[pm], minlength=case.Q)
        rows = np.nonzero(urow & ((pos_len > 0) | (cname == "overall")))[0]
        if cname != "overall" and rows.size == 0:
            assert cname not in res
            continue
        got = res[cname].cpu().numpy()
        assert got.shape == (rows.size, case.k + 1), (cname, got.shape, rows.size)
        assert (got[:, -1] == pos_len[rows]).all(), cname
        srt = -np.sort(-o.order_key(seg[rows]), axis=1)
        with np.errstate(invalid="ignore"):
            clear = (srt[:, case.k - 1] - srt[:, case.k]) > 1e-5 * scale
        _, idx = o.topk(seg[rows], case.k)
        pos_set = set(zip(pos_u[pm].tolist(), pos_i[pm].tolist()))
        want = np.array([[1 if (int(u), int(i)) in pos_set else 0 for i in idx[r]] for r, u in enumerate(rows)], dtype=np.int32)
        finite = np.isfinite(np.take_along_axis(seg[rows], idx, axis=1))
        want = np.where(finite, want, 0)                        # -inf slots (fewer than k candidates) are empty: no hit
        assert (got[clear][:, :-1] == want[clear]).all(), cname
        acc = ev.collectors[cname].get_data_struct()["rec.topk"].numpy()
        assert (acc == got).all(), cname


@pytest.mark.parametrize("name", ["bpr_lsh_ml100k", "directau_slsh", "bpr_mean"])
def test_sampled_negative_eval_vs_reference_golden(name, G):
    """SURVEY §8f row 4 (eval side): `model.pair_topk` — CSR of the (user, item) pairs, one embed of the candidates, fp32
    dot products, per-row selection — against the reference's neg_sample_batch_eval matrix (tests/golden/sampled_eval.npz:
    model.predict scattered into [users, N] of -inf) for the all / old-items / new-items segments: tie-aware sets, values
    rtol 1e-5, (score desc, id asc) order, (-inf, -1) where a row has fewer than k candidates (rows with 5 + positives
    candidates and duplicate pairs are in the fixture).  Then `InductiveEvaluator.neg_sample_batch_eval`: its seven
    collectors equal `topk_hits_collectors` on those lists."""
    import oov_b200
    from oov_b200 import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sampled_eval.npz"))
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    cfg, emb, model = G.build_retrieval(case, inp)
    rows, us, its = g[f"{name}.rows"], g[f"{name}.users"], g[f"{name}.items"]
    pos_u, pos_i = g[f"{name}.pos_u"], g[f"{name}.pos_i"]
    n_rows = int(rows.max()) + 1
    dense = g[f"{name}.scores_dense"]
    segs = ((0, 1 << 62), (0, case.n_old_items), (case.n_old_items, 1 << 62))
    # pairs in a shuffled order: the CSR build sorts them
    perm = np.random.default_rng(3).permutation(rows.size)
    out = model.pair_topk(G.t(rows[perm]), G.t(us[perm]), G.t(its[perm]), n_rows, case.k, segs=segs)
    lists = []
    for (s_, i_), (lo, hi) in zip(out, segs):
        seg = dense.copy()
        seg[:, :lo] = -np.inf
        seg[:, min(hi, case.n_all_items):] = -np.inf
        pu.sampled_rows_match(seg, s_.cpu().numpy(), i_.cpu().numpy(), case.k)
        lists.append(i_)
    # all-items list == merge of the two segment lists
    ms, mi = ops.topk_merge(torch.stack([out[1][0], out[2][0]]), torch.stack([out[1][1], out[2][1]]))
    assert torch.equal(mi, out[0][1]) and torch.equal(ms, out[0][0])
    ev = oov_b200.InductiveEvaluator(model, cfg, case.n_old_users, case.n_old_items)
    batch = ({"user_id": G.t(us), "item_id": G.t(its)}, G.t(rows), G.t(pos_u), G.t(pos_i))
    res = ev.neg_sample_batch_eval(batch)
    users_of_row = torch.zeros(n_rows, dtype=torch.int64, device=G.DEV)
    users_of_row[G.t(rows)] = G.t(us)
    rp, pc = ops.pairs_to_csr(G.t(pos_u), G.t(pos_i), n_rows)
    want = ops.topk_hits_collectors(lists[0], lists[1], lists[2], users_of_row, case.n_old_users, case.n_old_items, rp, pc)
    assert torch.equal(res.device_rows, want)
    overall = res["overall"].numpy()
    assert overall.shape == (n_rows, case.k + 1)
    pos = set(zip(pos_u.tolist(), pos_i.tolist()))
    idx_all = lists[0].cpu().numpy()
    hits = np.array([[1 if (r, int(i)) in pos else 0 for i in idx_all[r]] for r in range(n_rows)])
    assert (overall[:, :-1] == hits).all() and (overall[:, -1] == np.bincount(pos_u, minlength=n_rows)).all()
    assert hits.sum() > 0
