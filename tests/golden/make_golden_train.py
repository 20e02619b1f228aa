"""Golden fixtures of the training-mode OOV path: the UNMODIFIED reference models (BPR / DirectAU with the lsh / slsh /
zero embedders) in `set_oov_train()` mode, `calculate_loss` on a padded batch, `loss.backward()` — the loss and the
gradients of the four tables.  Authoring container only (needs /root/reference):
    python tests/golden/make_golden_train.py   ->   tests/golden/train_oov.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import cases  # noqa: E402
from oracle import refshim  # noqa: E402
from make_golden import base_config, interaction, ns, T  # noqa: E402


def run(case: cases.RetrievalCase) -> dict:
    inp = cases.retrieval_inputs(case)
    cfg = base_config(case, case.embedder, user_oov_buckets=case.B_user, item_oov_buckets=case.B_item)
    ds = refshim.RefDataset(case.n_old_users, case.n_old_items, interaction("user_id", inp["user_cols"]), interaction("item_id", inp["item_cols"]))
    emb = ns.get_inductive.get_inductive_embedder(cfg, ds, mode=f"golden-train-{case.name}", user_num=case.n_old_users, item_num=case.n_old_items)
    if case.embedder in ("lsh", "slsh"):
        emb.user_lsh.uniform_planes[0].data.copy_(T(inp["user_planes"]))
        emb.item_lsh.uniform_planes[0].data.copy_(T(inp["item_planes"]))
    model = (ns.BPR if case.model == "BPR" else ns.DirectAU)(cfg, ds, inductive_mapper=None, inductive_embedder=emb)
    with torch.no_grad():
        model.user_embedding.weight.copy_(T(inp["user_table"]))
        model.item_embedding.weight.copy_(T(inp["item_table"]))
        model.user_oov_buckets.weight.copy_(T(inp["user_oov"]))
        model.item_oov_buckets.weight.copy_(T(inp["item_oov"]))
    model.train()
    model.set_oov_train()                                           # abstract_recommender.py:147-154
    b = cases.train_batch(case)
    inter = ns.Interaction({"user_id": T(b["users"].copy()), "item_id": T(b["pos"].copy()), "neg_item_id": T(b["neg"].copy())})
    loss = model.calculate_loss(inter)                              # bpr.py:132-144 / directau.py:87-99
    loss.backward()
    out = {"loss": loss.detach().numpy()}
    for nm in ("user_embedding", "item_embedding", "user_oov_buckets", "item_oov_buckets"):
        gr = getattr(model, nm).weight.grad
        out[f"grad_{nm}"] = np.zeros(tuple(getattr(model, nm).weight.shape), np.float32) if gr is None else gr.numpy()
    return out


if __name__ == "__main__":
    res = {}
    for name in cases.TRAIN_CASES:
        for k, v in run(cases.CASES[name]).items():
            res[f"{name}.{k}"] = v
        print(name, "loss", res[f"{name}.loss"], {k.split(".")[1]: (float(np.nanmax(np.abs(v))), int(np.isnan(v).sum())) for k, v in res.items()
                                                 if k.startswith(name + ".grad")})
    np.savez_compressed(os.path.join(HERE, "train_oov.npz"), **res)
