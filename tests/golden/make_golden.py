"""Generate the golden fixtures by running the UNMODIFIED reference on seeded inputs.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Inputs come from `tests/golden/cases.py` (seeded numpy); this script feeds them to
the reference classes (imported through `oracle/refshim.py`, which only stubs
missing pip packages) and stores the reference's OUTPUTS in `tests/golden/*.npz`.
The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import refshim  # noqa: E402

ns = refshim.load()
T = torch.from_numpy


def interaction(id_field, cols):
    n = cols[0].shape[0]
    d = {id_field: torch.arange(n)}
    for i, c in enumerate(cols):
        d[f"f{i}"] = T(c)
    return ns.Interaction(d)


def base_config(case_like, embedder, **extra):
    cfg = refshim.RefConfig(
        USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cpu",
        embedding_size=case_like.D, add_oov_buckets=True,
        inductive_embedder=embedder, oov_prime_pad=cases.OOV_PRIME_PAD,
        oov_normalization_type=getattr(case_like, "normalization", "per-feature"),
        dhe_num_hashes=128, dhe_layer_size=512, gamma=1.0, oov_freeze_embedding=False,
        metrics=["Recall", "NDCG", "Hit"], topk=[10, 20], eval_args={"mode": "full"},
        model_eval_type="retrieval",
    )
    cfg.update(extra)
    return cfg


def sparse_near_zero(r: np.ndarray, thr: float = 1e-4):
    rows, cols = np.nonzero(np.abs(r) < thr)
    return rows.astype(np.int32), cols.astype(np.int32), r[rows, cols].astype(np.float32)


def run_retrieval(case: cases.RetrievalCase) -> dict:
    inp = cases.retrieval_inputs(case)
    uf = interaction("user_id", inp["user_cols"])
    itf = interaction("item_id", inp["item_cols"])
    cfg = base_config(case, case.embedder, user_oov_buckets=case.B_user, item_oov_buckets=case.B_item)
    ds = refshim.RefDataset(case.n_old_users, case.n_old_items, uf, itf)
    # factory path, inductive mode, ORIGINAL counts (src/perform_hashing.py:141-149)
    emb = ns.get_inductive.get_inductive_embedder(cfg, ds, mode=f"golden-{case.name}",
                                                  user_num=case.n_old_users, item_num=case.n_old_items)
    out = {}
    if case.embedder in ("lsh", "slsh"):
        emb.user_lsh.uniform_planes[0].data.copy_(T(inp["user_planes"]))
        emb.item_lsh.uniform_planes[0].data.copy_(T(inp["item_planes"]))
        out["user_feature_mat"] = emb.user_feature_mat.numpy()
        out["item_feature_mat"] = emb.item_feature_mat.numpy()
    model_cls = ns.BPR if case.model == "BPR" else ns.DirectAU
    model = model_cls(cfg, ds, inductive_mapper=None, inductive_embedder=emb).eval()
    with torch.no_grad():
        model.user_embedding.weight.copy_(T(inp["user_table"]))
        model.item_embedding.weight.copy_(T(inp["item_table"]))
        model.user_oov_buckets.weight.copy_(T(inp["user_oov"]))
        model.item_oov_buckets.weight.copy_(T(inp["item_oov"]))
        out["state_dict_keys"] = np.array(sorted(model.state_dict().keys()))

        oov_users = torch.arange(case.n_old_users, case.n_all_users)
        oov_items = torch.arange(case.n_old_items, case.n_all_items)
        if case.embedder in ("lsh", "slsh"):
            hu = emb._hash_users(oov_users.clone())
            hi = emb._hash_items(oov_items.clone())
            # the projections exactly as torch_hash.py:56 computes them (for tie bookkeeping)
            ru = (emb.user_feature_mat[oov_users] @ emb.user_lsh.uniform_planes[0].data.T).numpy()
            ri = (emb.item_feature_mat[oov_items] @ emb.item_lsh.uniform_planes[0].data.T).numpy()
            for nm, r in (("user", ru), ("item", ri)):
                rr, cc, vv = sparse_near_zero(r)
                out[f"{nm}_near_rows"], out[f"{nm}_near_cols"], out[f"{nm}_near_vals"] = rr, cc, vv
            if case.embedder == "lsh":
                out["user_bits"] = np.packbits(hu.numpy().astype(np.uint8), axis=1, bitorder="little")
                out["item_bits"] = np.packbits(hi.numpy().astype(np.uint8), axis=1, bitorder="little")
            else:
                out["user_bucket_ids"] = hu.numpy().astype(np.int64)
                out["item_bucket_ids"] = hi.numpy().astype(np.int64)
        out["oov_user_emb"] = emb.embed_user_ids(oov_users.clone(), model).numpy()
        out["oov_item_emb"] = emb.embed_item_ids(oov_items.clone(), model).numpy()
        if case.embedder == "lsh":
            # training-mode ids carry the prime pad and are de-padded in place (lsh_embedder.py:153-155)
            emb.set_train()
            padded = oov_items[:16].clone() + cases.OOV_PRIME_PAD
            out["oov_item_emb_train16"] = emb.embed_item_ids(padded, model).numpy()
            out["padded_after"] = padded.numpy()
            emb.set_eval()

        item_range = torch.arange(case.n_all_items)
        users = T(inp["users"])
        out["user_e"] = model.get_user_embedding(users.clone()).numpy()
        out["all_item_e"] = model.get_item_embedding(item_range.clone()).numpy()
        scores = model.ind_full_sort_predict(ns.Interaction({"user_id": users.clone()}), item_range)
        scores = scores.view(-1, case.n_all_items)
        out["scores_raw"] = scores.numpy().copy()
        # inductive/evaluator.py:91-94
        scores[:, 0] = -np.inf
        scores[(T(inp["hist_u"]), T(inp["hist_i"]))] = -np.inf
        vals, idx = torch.topk(scores, case.k, dim=-1)           # evaluator/collector.py:153
        out["topk_vals"] = vals.numpy()
        out["topk_idx"] = idx.numpy().astype(np.int64)

        # --- the collectors themselves ('rec.topk' = [hits | pos_len], collector.py:157-166) ---
        from recbole.evaluator.collector import Collector
        from recbole.inductive.filtered_collector import FilteredCollector
        from recbole.inductive.collector_filter import FastUserItemCollectorFilter
        ccfg = base_config(case, case.embedder, topk=[min(10, case.k), case.k])
        inter = ns.Interaction({"user_id": users.clone(), "item_id": torch.zeros_like(users)})
        pos_u, pos_i = T(inp["pos_u"]), T(inp["pos_i"])
        col = Collector(ccfg)
        col.eval_batch_collect(scores.clone(), inter, pos_u, pos_i)
        out["collector_overall"] = col.get_data_struct().get("rec.topk").numpy()
        for nm, ru_, ri_ in (("old_users", True, None), ("new_users", False, None),
                             ("old_old", True, True), ("old_new", True, False),
                             ("new_old", False, True), ("new_new", False, False)):
            flt = FastUserItemCollectorFilter(case.n_old_users, case.n_old_items,
                                              return_old_users=ru_, return_old_items=ri_)
            fc = FilteredCollector(ccfg, flt, nm)
            torch.manual_seed(0)
            ok = fc.eval_batch_collect(scores.clone(), inter, pos_u.clone(), pos_i.clone())
            if ok:
                out[f"collector_{nm}"] = fc.get_data_struct().get("rec.topk").numpy()
                out[f"collector_{nm}_rows"] = flt.last_users.unique(sorted=True).numpy()
    return out


def run_dhe(case: cases.DheCase) -> dict:
    keys = cases.dhe_keys(case.seed, case.n_hashes)
    ids = cases.dhe_ids(case)
    ws, bs = cases.dhe_weights(case)
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                       # dh_embedder.py:52,103-105: ./hash_keys relative to the CWD
        try:
            os.makedirs("hash_keys")
            import json
            with open(f"hash_keys/{case.n_hashes}.hashes", "w") as f:
                json.dump([k.hex() for k in keys], f)
            n = 8
            feats_u = ns.Interaction({"user_id": torch.arange(n), "f0": torch.ones(n, 2)})
            feats_i = ns.Interaction({"item_id": torch.arange(n), "f0": torch.ones(n, 2)})
            cfg = base_config(case, "dhe", user_oov_buckets=4, item_oov_buckets=4, dhe_num_hashes=case.n_hashes)
            ds = refshim.RefDataset(4, 4, feats_u, feats_i)
            emb = ns.get_inductive.get_inductive_embedder(cfg, ds, mode=f"golden-{case.name}", user_num=4, item_num=4)
            assert [bytes(k) for k in emb.hash_keys] == keys
            out["state_dict_keys"] = np.array(sorted(emb.state_dict().keys()))
            with torch.no_grad():
                for l, li in enumerate((0, 2, 4, 6)):
                    emb.item_hash_net[li].weight.copy_(T(ws[l]))
                    emb.item_hash_net[li].bias.copy_(T(bs[l]))
                    emb.user_hash_net[li].weight.copy_(T(ws[l][::-1].copy()))     # a different net for users
                    emb.user_hash_net[li].bias.copy_(T(bs[l][::-1].copy()))
                h = emb._hash_ids(T(ids))
                out["hashes"] = h.numpy().astype(np.uint32)
                assert (h.numpy() == out["hashes"]).all()
                out["item_emb"] = emb.embed_item_ids(T(ids), None).numpy()
                out["user_emb"] = emb.embed_user_ids(T(ids), None).numpy()
                # pre-sigmoid logits: lets the tests bound the error where sigmoid saturates
                z = emb.item_hash_net[:-1](h.float())
                out["item_logits"] = z.numpy()
        finally:
            os.chdir(cwd)
    return out


def _make_context_model(case: cases.ContextCase, inp, emb, D, table, user_oov, item_oov, first_order=False):
    """Duck-typed `self` for the reference's unbound embed_token_fields (constructing a
    full DCNV2 needs the whole RecBole data pipeline).  isinstance checks in
    mean_embedder.py:55-60 need a real model class, so allocate DCNV2 without __init__."""
    from recbole.model.context_aware_recommender.dcnv2 import DCNV2
    from recbole.model.layers import FMEmbedding, InductiveFMFirstOrderLinear
    cls = InductiveFMFirstOrderLinear if first_order else DCNV2
    m = cls.__new__(cls)
    torch.nn.Module.__init__(m)
    m.n_users, m.n_items = case.n_old_users, case.n_old_items
    m.inductive_mapper, m.inductive_embedder = None, emb
    m.token_embedding_table = FMEmbedding(inp["dims"].tolist(), inp["offsets"], D)
    m.token_field_offsets = inp["offsets"]
    m.user_oov_buckets = torch.nn.Embedding(case.B, D)
    m.item_oov_buckets = torch.nn.Embedding(case.B, D)
    with torch.no_grad():
        m.token_embedding_table.embedding.weight.copy_(T(table))
        m.user_oov_buckets.weight.copy_(T(user_oov))
        m.item_oov_buckets.weight.copy_(T(item_oov))
    return m


def run_context(case: cases.ContextCase) -> dict:
    from recbole.model.abstract_recommender import InductiveContextRecommender
    from recbole.model.layers import InductiveFMFirstOrderLinear
    inp = cases.context_inputs(case)
    uf = interaction("user_id", inp["user_cols"])
    itf = interaction("item_id", inp["item_cols"])
    out = {}
    for tag, D, table, uo, io, pu, pi in (
            ("", case.D, inp["table"], inp["user_oov"], inp["item_oov"], inp["user_planes"], inp["item_planes"]),
            ("1", 1, inp["table1"], inp["user_oov1"], inp["item_oov1"], inp["user_planes1"], inp["item_planes1"])):
        cfg = base_config(case, case.embedder, user_oov_buckets=case.B, item_oov_buckets=case.B, embedding_size=D)
        ds = refshim.RefDataset(case.n_old_users, case.n_old_items, uf, itf)
        emb = ns.get_inductive.get_inductive_embedder(cfg, ds, mode=f"golden-{case.name}{tag}",
                                                      user_num=case.n_old_users, item_num=case.n_old_items,
                                                      embedding_size=D)
        if case.embedder in ("lsh", "slsh"):
            emb.user_lsh.uniform_planes[0].data.copy_(T(pu))
            emb.item_lsh.uniform_planes[0].data.copy_(T(pi))
        m = _make_context_model(case, inp, emb, D, table, uo, io, first_order=(tag == "1"))
        with torch.no_grad():
            if tag == "":
                e = InductiveContextRecommender.embed_token_fields(m, T(inp["tokens"]).clone())
                out["token_embedding"] = e.numpy()
            else:
                e = InductiveFMFirstOrderLinear.embed_token_fields(m, T(inp["tokens"]).clone(), 0, 1)
                out["first_order_sum"] = e.numpy()
    return out


def run_mapper() -> dict:
    out = {}
    g = cases.rng(77)
    ids = np.concatenate([np.arange(0, 64), g.integers(0, 1 << 40, size=192)]).astype(np.int64)
    out["ids"] = ids
    n = 8
    fu = ns.Interaction({"user_id": torch.arange(n)})
    for fn in ("mod", "fast", "3round", "64bit"):
        for nb in (1000, 7):
            mp = ns.RandomOOVInductiveMapper(fu, fu, 50, 50, nb, nb, 8, "cpu", cases.OOV_PRIME_PAD, fn)
            out[f"{fn}_{nb}"] = mp.map_item_ids(T(ids).clone()).numpy()
    return out


def main():
    written = []
    for name, case in cases.CASES.items():
        out = run_retrieval(case)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        written.append(name)
    for name, case in cases.DHE_CASES.items():
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **run_dhe(case))
        written.append(name)
    for name, case in cases.CONTEXT_CASES.items():
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **run_context(case))
        written.append(name)
    np.savez_compressed(os.path.join(HERE, "random_mapper.npz"), **run_mapper())
    written.append("random_mapper")
    for w in written:
        p = os.path.join(HERE, f"{w}.npz")
        print(f"{w:28s} {os.path.getsize(p) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
