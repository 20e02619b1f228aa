"""Pin the oracle (numpy/C restatement) against the golden fixtures, i.e. against the
outputs of the UNMODIFIED reference run through oracle/refshim.py by
tests/golden/make_golden.py.  CPU only."""
import os
import numpy as np
import pytest

import cases
import parity_util as pu
from oracle import oracle as o


@pytest.mark.parametrize("name", list(cases.CASES))
def test_retrieval_case(name):
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    g = pu.load_golden(name)
    out = pu.oracle_retrieval(case, inp)

    skip_u = skip_i = np.zeros(0, np.int64)
    if case.embedder in ("lsh", "slsh"):
        pu.assert_close(out["user_feature_mat"], g["user_feature_mat"], rtol=2e-6, atol=1e-7, what="user_feature_mat")
        pu.assert_close(out["item_feature_mat"], g["item_feature_mat"], rtol=2e-6, atol=1e-7, what="item_feature_mat")
    if case.embedder == "lsh":
        for side, B in (("user", case.B_user), ("item", case.B_item)):
            want = pu.unpack_bits(g[f"{side}_bits"], B)
            ties = pu.tie_positions(g[f"{side}_near_rows"], g[f"{side}_near_cols"], g[f"{side}_near_vals"])
            pu.check_bits(out[f"{side}_bits"], want, ties)
        skip_u = pu.rows_with_bit_diffs(out["user_bits"], pu.unpack_bits(g["user_bits"], case.B_user))
        skip_i = pu.rows_with_bit_diffs(out["item_bits"], pu.unpack_bits(g["item_bits"], case.B_item))
    if case.embedder == "slsh":
        for side in ("user", "item"):
            got, want = out[f"{side}_bucket_ids"], g[f"{side}_bucket_ids"]
            diff = np.nonzero(got != want)[0]
            tie_rows = set(g[f"{side}_near_rows"][np.abs(g[f"{side}_near_vals"]) < pu.TIE_EPS].tolist())
            assert set(diff.tolist()) <= tie_rows
        skip_u = np.nonzero(out["user_bucket_ids"] != g["user_bucket_ids"])[0]
        skip_i = np.nonzero(out["item_bucket_ids"] != g["item_bucket_ids"])[0]

    pu.assert_close(out["oov_user_emb"], g["oov_user_emb"], skip_rows=skip_u, what="oov_user_emb")
    pu.assert_close(out["oov_item_emb"], g["oov_item_emb"], skip_rows=skip_i, what="oov_item_emb")
    if len(skip_u) == 0 and len(skip_i) == 0:
        pu.assert_close(out["user_e"], g["user_e"], what="user_e")
        pu.assert_close(out["all_item_e"], g["all_item_e"], what="all_item_e")
        # scores: sums of D products of magnitude |u||v| -> atol scaled to that magnitude
        scale = float(np.nanmax(np.abs(g["scores_raw"][np.isfinite(g["scores_raw"])]))) if np.isfinite(g["scores_raw"]).any() else 1.0
        pu.assert_close(out["scores_raw"], g["scores_raw"], rtol=1e-5, atol=1e-5 * scale, what="scores_raw")
        ok, msg = o.topk_sets_match(out["scores_masked"], g["topk_idx"], case.k, rtol=1e-5, atol=1e-6 * scale)
        assert ok, msg
        # collector 'rec.topk' = [hits | pos_len] (collector.py:157-166); rows with a clear k-th gap only
        mine = o.collector_hits(out["topk_idx"], inp["pos_u"], inp["pos_i"], case.n_all_items)
        srt = -np.sort(-o.order_key(out["scores_masked"]), axis=1)
        with np.errstate(invalid="ignore"):
            clear = (srt[:, case.k - 1] - srt[:, case.k]) > 1e-5 * scale
        assert (mine[:, -1] == g["collector_overall"][:, -1]).all()
        assert (mine[clear].sum(1) == g["collector_overall"][clear].sum(1)).all()
        assert clear.sum() >= 1 or case.name == "bpr_lsh_tinybuckets"


def test_lsh_training_mode_depads_ids():
    """lsh_embedder.py:173-175: ids >= prime_pad are de-padded (in place) in training mode."""
    case = cases.CASES["bpr_lsh_ml100k"]
    inp = cases.retrieval_inputs(case)
    g = pu.load_golden(case.name)
    ifm = o.feature_matrix(inp["item_cols"], case.normalization)
    ids = np.arange(case.n_old_items, case.n_old_items + 16) + cases.OOV_PRIME_PAD
    got = o.lsh_embed(ifm, ids, inp["item_planes"], inp["item_oov"], training=True)
    pu.assert_close(got, g["oov_item_emb_train16"], what="train-mode emb")
    assert (g["padded_after"] == ids - cases.OOV_PRIME_PAD).all()      # the reference mutates its input


def test_lsh_nan_rows_exist_and_match():
    g = pu.load_golden("bpr_lsh_tinybuckets")
    case = cases.CASES["bpr_lsh_tinybuckets"]
    out = pu.oracle_retrieval(case, cases.retrieval_inputs(case))
    nan_rows = np.isnan(g["oov_item_emb"]).any(axis=1)
    assert nan_rows.sum() > 0
    assert (np.isnan(out["oov_item_emb"]).any(axis=1) == nan_rows).all()


@pytest.mark.parametrize("name", list(cases.DHE_CASES))
def test_dhe_case(name):
    case = cases.DHE_CASES[name]
    g = pu.load_golden(name)
    keys = o.keys_to_array(cases.dhe_keys(case.seed, case.n_hashes))
    ids = cases.dhe_ids(case)
    ws, bs = cases.dhe_weights(case)
    h = o.dhe_hashes(ids, keys)
    assert (h == g["hashes"]).all()                                # bit-exact hash ids
    emb = o.dhe_mlp(h, ws, bs)
    if case.w1_scale == 1.0:
        # default init saturates: logits ~1e6, outputs exactly 0/1 in the reference
        sat = np.abs(g["item_logits"]) > 120
        assert sat.mean() > 0.99
        assert (emb[sat] == g["item_emb"][sat]).all()
    else:
        pu.assert_close(emb, g["item_emb"], rtol=1e-5, atol=1e-6, what="dhe item_emb")
    emb_u = o.dhe_mlp(h, [w[::-1] for w in ws], [b[::-1] for b in bs])
    if case.w1_scale != 1.0:
        pu.assert_close(emb_u, g["user_emb"], rtol=1e-5, atol=1e-6, what="dhe user_emb")
    want_keys = {f"{side}_hash_net.{i}.{p}" for side in ("user", "item") for i in (0, 2, 4, 6) for p in ("weight", "bias")}
    assert want_keys <= set(g["state_dict_keys"].tolist())


@pytest.mark.parametrize("name", list(cases.CONTEXT_CASES))
def test_context_case(name):
    case = cases.CONTEXT_CASES[name]
    inp = cases.context_inputs(case)
    g = pu.load_golden(name)
    ufm = o.feature_matrix(inp["user_cols"], "per-feature")
    ifm = o.feature_matrix(inp["item_cols"], "per-feature")

    def embedders(pu_, pi_, uo, io, table, d):
        off = inp["offsets"]
        if case.embedder == "lsh":
            return (lambda ids: o.lsh_embed(ufm, ids, pu_, uo)), (lambda ids: o.lsh_embed(ifm, ids, pi_, io))
        if case.embedder == "slsh":
            return (lambda ids: o.slsh_embed(ufm, ids, pu_, uo)), (lambda ids: o.slsh_embed(ifm, ids, pi_, io))
        if case.embedder == "mean":     # mean_embedder.py:57-60, 72-85: slices of the token table
            return (lambda ids: o.mean_embed(table[off[0]:off[1]], len(ids))), \
                   (lambda ids: o.mean_embed(table[off[1]:off[2]], len(ids)))
        return (lambda ids: o.zero_embed(len(ids), d)), (lambda ids: o.zero_embed(len(ids), d))

    eu, ei = embedders(inp["user_planes"], inp["item_planes"], inp["user_oov"], inp["item_oov"], inp["table"], case.D)
    got = o.embed_token_fields(inp["tokens"], inp["offsets"], inp["table"], case.n_old_users, case.n_old_items, eu, ei)
    pu.assert_close(got.reshape(got.shape[0], -1), g["token_embedding"].reshape(got.shape[0], -1), what="token_embedding")
    eu1, ei1 = embedders(inp["user_planes1"], inp["item_planes1"], inp["user_oov1"], inp["item_oov1"], inp["table1"], 1)
    got1 = o.first_order_token_sum(inp["tokens"], inp["offsets"], inp["table1"], case.n_old_users, case.n_old_items, eu1, ei1)
    pu.assert_close(got1.reshape(-1, 1), g["first_order_sum"].reshape(-1, 1), rtol=1e-5, atol=2e-6, what="first_order_sum")


def test_random_mapper():
    g = pu.load_golden("random_mapper")
    ids = g["ids"]
    for fn in ("mod", "fast", "3round", "64bit"):
        for nb in (1000, 7):
            assert (o.map_ids(ids, 50, nb, fn) == g[f"{fn}_{nb}"]).all(), (fn, nb)


def test_topk_helpers():
    s = np.array([[1.0, 3.0, 3.0, 2.0, -np.inf], [np.nan, 0.0, 5.0, np.nan, 1.0]], np.float32)
    vals, idx = o.topk(s, 2)
    assert idx.tolist() == [[1, 2], [0, 3]]                         # score desc, index asc; NaN ranks first
    assert o.topk_sets_match(s, np.array([[2, 1], [3, 0]]), 2)[0]
    assert not o.topk_sets_match(s, np.array([[1, 3], [3, 0]]), 2)[0]
    assert o.topk_sets_match(s, np.array([[1, 2, 3], [0, 3, 2]]), 3)[0]
    cs = np.stack([vals, vals - 1])
    ci = np.stack([idx, idx + 10])
    mv, mi = o.merge_topk(cs, ci, 2)
    assert mi.tolist() == idx.tolist()


def _dcnv2_case(g, name):
    pre = name + "."
    cw = [g[pre + f"cross_w{l}"] for l in range(3)]
    cb = [g[pre + f"cross_b{l}"] for l in range(3)]
    layers = []
    l = 0
    while pre + f"mlp_w{l}" in g.files:
        layers.append(dict(w=g[pre + f"mlp_w{l}"], b=g[pre + f"mlp_b{l}"], bn_mean=g[pre + f"bn_mean{l}"], bn_var=g[pre + f"bn_var{l}"],
                           bn_gamma=g[pre + f"bn_gamma{l}"], bn_beta=g[pre + f"bn_beta{l}"], bn_eps=float(g[pre + f"bn_eps{l}"])))
        l += 1
    return g[pre + "x0"], cw, cb, layers, g[pre + "pred_w"], g[pre + "pred_b"]


@pytest.mark.parametrize("structure", ["stacked", "parallel"])
def test_oracle_dcnv2_tower_vs_reference(structure):
    """oracle.dcnv2_cross / mlp_bn_relu / dcnv2_forward against the reference's DCNV2.cross_network + MLPLayers(eval) +
    predict layer (tests/golden/make_golden_dcnv2.py), fp32 rtol 1e-5; BatchNorm folding is exact to fp32 rounding."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dcnv2_tower.npz"))
    x0, cw, cb, layers, pw, pb = _dcnv2_case(g, structure)
    pre = structure + "."
    cross = o.dcnv2_cross(x0, cw, cb)
    np.testing.assert_allclose(cross, g[pre + "cross_out"], rtol=1e-5, atol=1e-5)
    deep_in = cross if structure == "stacked" else x0
    np.testing.assert_allclose(o.mlp_bn_relu(deep_in, layers), g[pre + "deep_out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(o.dcnv2_forward(x0, cw, cb, layers, pw, pb, structure), g[pre + "out"], rtol=1e-5, atol=1e-6)
    # folded BatchNorm == BatchNorm
    h = deep_in
    for L in layers:
        w, b = o.fold_bn(L)
        h = np.maximum(h @ w.T + b, 0)
    np.testing.assert_allclose(h, g[pre + "deep_out"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["bpr_lsh_ml100k", "directau_slsh", "bpr_mean"])
def test_oracle_sampled_negative_eval_vs_reference(name):
    """oracle.pair_scores / neg_sample_scores against the reference's model.predict on (user, item) pairs and
    InductiveEvaluator.neg_sample_batch_eval (tests/golden/make_golden_sampled.py): origin scores rtol 1e-5, the dense
    [users, N] matrix equal where finite and -inf elsewhere, per-segment top-k sets tie-aware."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sampled_eval.npz"))
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    ora = pu.oracle_retrieval(case, inp)
    rows, us, its = g[f"{name}.rows"], g[f"{name}.users"], g[f"{name}.items"]
    origin = o.pair_scores(ora["all_user_e"][us], ora["all_item_e"][its], normalize=case.model == "DirectAU")
    pu.assert_close(origin, g[f"{name}.origin_scores"], what="origin scores")
    n_rows = int(rows.max()) + 1
    dense = o.neg_sample_scores(origin, rows, its, n_rows, case.n_all_items)
    want = g[f"{name}.scores_dense"]
    assert (np.isneginf(dense) == np.isneginf(want)).all()
    fin = ~np.isneginf(want)
    pu.assert_close(dense[fin], want[fin], what="dense scores")
    for nm, lo, hi in (("all", 0, case.n_all_items), ("old", 0, case.n_old_items), ("new", case.n_old_items, case.n_all_items)):
        seg = want.copy()
        seg[:, :lo] = -np.inf
        seg[:, hi:] = -np.inf
        vals, idx = o.topk(np.where(np.isneginf(seg), seg, dense), case.k)
        pu.sampled_rows_match(seg, np.where(np.isfinite(vals) | np.isnan(vals), vals, -np.inf),
                              np.where(np.isneginf(vals), -1, idx), case.k)


def _widedeep_case(g, name):
    pre = name + "."
    ws, bs = [], []
    l = 0
    while pre + f"mlp_w{l}" in g.files:
        ws.append(g[pre + f"mlp_w{l}"]); bs.append(g[pre + f"mlp_b{l}"])
        l += 1
    return g[pre + "emb"], g[pre + "fm"], ws, bs, g[pre + "pred_w"], g[pre + "pred_b"]


@pytest.mark.parametrize("name", ["default", "wide"])
def test_oracle_widedeep_head_vs_reference(name):
    """oracle.widedeep_forward against the reference's WideDeep.forward / predict (tests/golden/make_golden_widedeep.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "widedeep_head.npz"))
    emb, fm, ws, bs, pw, pb = _widedeep_case(g, name)
    logits = o.widedeep_forward(emb, fm, ws, bs, pw, pb)
    np.testing.assert_allclose(logits, g[name + ".logits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(o.sigmoid(logits), g[name + ".prob"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", list(cases.FEATNET_CASES))
def test_featnet_case(name):
    """fdhe / dnn (feat_dh_embedder.py:188-213, dnn_embedder.py:87-109): the oracle restatement against the outputs of
    the unmodified reference classes, eval and training mode (hashes of the padded id, features of the de-padded id)."""
    case = cases.FEATNET_CASES[name]
    inp = cases.featnet_inputs(case)
    g = np.load(os.path.join(pu.GOLDEN_DIR, "featnet.npz"), allow_pickle=False)
    keys = o.keys_to_array(cases.dhe_keys(case.seed, case.n_hashes)) if case.kind == "fdhe" else None
    for side in ("user", "item"):
        fm = o.featnet_feature_matrix(inp[f"{side}_cols"])
        pu.assert_close(fm, g[f"{name}.{side}_feature_mat"], rtol=2e-6, atol=1e-7, what="feature_mat")
        ws, bs = inp["nets"][side]
        pu.assert_close(o.fdhe_embed(inp["ids"], keys, fm, ws, bs), g[f"{name}.{side}_emb"], rtol=1e-5, atol=1e-6, what=f"{side}_emb")
        pu.assert_close(o.fdhe_embed(inp["ids_train"], keys, fm, ws, bs, training=True), g[f"{name}.{side}_emb_train"],
                        rtol=1e-5, atol=1e-6, what=f"{side}_emb_train")
    if case.kind == "fdhe":        # the pad changes the hashes, so the two modes must differ
        assert np.abs(g[f"{name}.item_emb_train"] - g[f"{name}.item_emb"]).max() > 1e-3
    want_keys = {f"{side}_hash_net.{i}.{p}" for side in ("user", "item") for i in (0, 2, 4, 6) for p in ("weight", "bias")}
    assert want_keys <= set(g[f"{name}.state_dict_keys"].tolist())


def _xdeepfm_case(g, name):
    pre = name + "."

    def seq(stem):
        out, l = [], 0
        while pre + f"{stem}{l}" in g.files:
            out.append(g[pre + f"{stem}{l}"])
            l += 1
        return out

    conv_w = [w[:, :, 0] for w in seq("conv_w")]
    return dict(emb=g[pre + "emb"], fm=g[pre + "fm"], conv_w=conv_w, conv_b=seq("conv_b"), lin_w=g[pre + "lin_w"], lin_b=g[pre + "lin_b"],
                mlp_w=seq("mlp_w"), mlp_b=seq("mlp_b"), direct=bool(int(g[pre + "direct"])))


@pytest.mark.parametrize("name", ["default", "direct", "odd"])
def test_oracle_xdeepfm_head_vs_reference(name):
    """oracle.xdeepfm_cin / xdeepfm_forward against the reference's xDeepFM.compressed_interaction_network / forward /
    predict (tests/golden/make_golden_xdeepfm.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "xdeepfm_head.npz"))
    c = _xdeepfm_case(g, name)
    cin = o.xdeepfm_cin(c["emb"], c["conv_w"], c["conv_b"], c["direct"])
    np.testing.assert_allclose(cin, g[name + ".cin"], rtol=2e-5, atol=2e-5)
    logits = o.xdeepfm_forward(c["emb"], c["fm"], c["conv_w"], c["conv_b"], c["lin_w"], c["lin_b"], c["mlp_w"], c["mlp_b"], c["direct"])
    np.testing.assert_allclose(logits, g[name + ".logits"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(o.sigmoid(logits), g[name + ".prob"], rtol=1e-5, atol=2e-6)
