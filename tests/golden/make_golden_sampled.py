"""Golden fixture for the sampled-negative evaluation path (eval side).

Runs the UNMODIFIED reference in the authoring container: BPR + lsh / DirectAU + mean models built exactly like
make_golden.run_retrieval builds them, `model.predict` on (user, item) pairs (bpr.py:146-149) and
`InductiveEvaluator.neg_sample_batch_eval` (inductive/evaluator.py:116-133: scatter into a [users, N] matrix of -inf),
then torch.topk per user and per item segment.  Writes tests/golden/sampled_eval.npz.

    python tests/golden/make_golden_sampled.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE]

import cases  # noqa: E402
import make_golden as mg  # noqa: E402  (imports the reference through oracle/refshim.py)

ns, T, refshim = mg.ns, mg.T, mg.refshim


def sampled_pairs(case, seed, n_rows=24, n_neg=37):
    """Pairs like a NegSampleEvalDataLoader batch: per batch user its positives followed by sampled negatives (ids over old
    AND new items, with a few duplicates), rows numbered 0 .. n_rows - 1."""
    g = cases.rng(seed)
    users = g.choice(np.arange(1, case.n_all_users), size=n_rows, replace=False).astype(np.int64)
    rows, us, its, pu, pi = [], [], [], [], []
    for r, u in enumerate(users):
        n_pos = int(g.integers(1, 4))
        pos = g.choice(np.arange(1, case.n_all_items), size=n_pos, replace=False)
        neg = g.integers(1, case.n_all_items, size=n_neg)
        if r % 5 == 0:
            neg[:3] = neg[3:6]                                   # duplicate pairs
        cand = np.concatenate([pos, neg]) if r % 7 else np.concatenate([pos, neg[:5]])     # some rows have fewer than k candidates
        rows += [r] * len(cand); us += [u] * len(cand); its += cand.tolist()
        pu += [r] * n_pos; pi += pos.tolist()
    return (np.asarray(rows, np.int64), np.asarray(us, np.int64), np.asarray(its, np.int64), np.asarray(pu, np.int64),
            np.asarray(pi, np.int64), users)


def run(case_name, seed):
    case = cases.CASES[case_name]
    inp = cases.retrieval_inputs(case)
    uf = mg.interaction("user_id", inp["user_cols"])
    itf = mg.interaction("item_id", inp["item_cols"])
    cfg = mg.base_config(case, case.embedder, user_oov_buckets=case.B_user, item_oov_buckets=case.B_item)
    ds = refshim.RefDataset(case.n_old_users, case.n_old_items, uf, itf)
    emb = ns.get_inductive.get_inductive_embedder(cfg, ds, mode=f"golden-sampled-{case.name}", user_num=case.n_old_users,
                                                  item_num=case.n_old_items)
    if case.embedder in ("lsh", "slsh"):
        emb.user_lsh.uniform_planes[0].data.copy_(T(inp["user_planes"]))
        emb.item_lsh.uniform_planes[0].data.copy_(T(inp["item_planes"]))
    model_cls = ns.BPR if case.model == "BPR" else ns.DirectAU
    model = model_cls(cfg, ds, inductive_mapper=None, inductive_embedder=emb).eval()
    out = {}
    with torch.no_grad():
        model.user_embedding.weight.copy_(T(inp["user_table"]))
        model.item_embedding.weight.copy_(T(inp["item_table"]))
        model.user_oov_buckets.weight.copy_(T(inp["user_oov"]))
        model.item_oov_buckets.weight.copy_(T(inp["item_oov"]))
        rows, us, its, pu, pi, users = sampled_pairs(case, seed)
        inter = ns.Interaction({"user_id": T(us), "item_id": T(its)})
        from recbole.inductive.evaluator import InductiveEvaluator
        from recbole.utils import EvaluatorType
        self = types.SimpleNamespace(model=model, device=torch.device("cpu"), test_batch_size=1 << 30,
                                     config={"eval_type": EvaluatorType.RANKING, "ITEM_ID_FIELD": "item_id"},
                                     tot_item_num=case.n_all_items)
        _, scores, _, _ = InductiveEvaluator.neg_sample_batch_eval(self, (inter, T(rows), T(pu), T(pi)))
        out["origin_scores"] = model.predict(ns.Interaction({"user_id": T(us), "item_id": T(its)})).numpy()
        out["scores_dense"] = scores.numpy()
        for nm, lo, hi in (("all", 0, case.n_all_items), ("old", 0, case.n_old_items), ("new", case.n_old_items, case.n_all_items)):
            sc = scores.clone()
            sc[:, :lo] = -np.inf
            sc[:, hi:] = -np.inf
            v, ix = torch.topk(sc, case.k, dim=-1)
            out[f"topk_vals_{nm}"] = v.numpy()
            out[f"topk_idx_{nm}"] = ix.numpy().astype(np.int64)
    out.update(rows=rows, users=us, items=its, pos_u=pu, pos_i=pi, batch_users=users)
    return out


def main():
    allout = {}
    for name, seed in (("bpr_lsh_ml100k", 501), ("directau_slsh", 502), ("bpr_mean", 503)):
        if name not in cases.CASES:
            continue
        for k_, v in run(name, seed).items():
            allout[f"{name}.{k_}"] = v
    np.savez_compressed(os.path.join(HERE, "sampled_eval.npz"), **allout)
    print("wrote sampled_eval.npz:", sorted({k.split('.')[0] for k in allout}))


if __name__ == "__main__":
    main()
