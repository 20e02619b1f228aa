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


def run_hashnet(case: cases.FeatNetCase) -> dict:
    """BPR + dhe / fdhe / dnn in training mode: loss, table gradients and the gradients of both hash nets."""
    import json
    import tempfile
    inp = cases.hashnet_train_inputs(case)
    keys = cases.dhe_keys(case.seed, case.n_hashes)
    n_old = inp["n_old"]
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            os.makedirs("hash_keys")
            with open(f"hash_keys/{case.n_hashes}.hashes", "w") as f:
                json.dump([k.hex() for k in keys], f)
            cfg = base_config(case, case.kind, user_oov_buckets=4, item_oov_buckets=4, dhe_num_hashes=case.n_hashes, dhe_layer_size=case.layer)
            ds = refshim.RefDataset(n_old, n_old, interaction("user_id", inp["user_cols"]), interaction("item_id", inp["item_cols"]))
            emb = ns.get_inductive.get_inductive_embedder(cfg, ds, mode=f"golden-train-{case.name}", user_num=n_old, item_num=n_old)
        finally:
            os.chdir(cwd)
    model = ns.BPR(cfg, ds, inductive_mapper=None, inductive_embedder=emb)
    with torch.no_grad():
        model.user_embedding.weight.copy_(T(inp["user_table"]))
        model.item_embedding.weight.copy_(T(inp["item_table"]))
        for side, net in (("user", emb.user_hash_net), ("item", emb.item_hash_net)):
            ws, bs = inp["nets"][side]
            for l, li in enumerate((0, 2, 4, 6)):
                net[li].weight.copy_(T(ws[l]))
                net[li].bias.copy_(T(bs[l]))
    model.train()
    model.set_oov_train()
    b = inp["batch"]
    inter = ns.Interaction({"user_id": T(b["users"].copy()), "item_id": T(b["pos"].copy()), "neg_item_id": T(b["neg"].copy())})
    loss = model.calculate_loss(inter)
    loss.backward()
    out = {"loss": loss.detach().numpy(), "grad_user_embedding": model.user_embedding.weight.grad.numpy(),
           "grad_item_embedding": model.item_embedding.weight.grad.numpy()}
    for side, net in (("user", emb.user_hash_net), ("item", emb.item_hash_net)):
        for li in (0, 2, 4, 6):
            out[f"grad_{side}_hash_net.{li}.weight"] = cases.grad_slice(net[li].weight.grad.numpy())
            out[f"grad_{side}_hash_net.{li}.bias"] = net[li].bias.grad.numpy()
    return out


if __name__ == "__main__":
    res = {}
    for name, case in cases.HASHNET_TRAIN_CASES.items():
        for k, v in run_hashnet(case).items():
            res[f"{name}.{k}"] = v
        print(name, "loss", res[f"{name}.loss"], {k.split("grad_")[1]: float(np.abs(v).max()) for k, v in res.items() if k.startswith(name + ".grad")})
    for name in cases.TRAIN_CASES:
        for k, v in run(cases.CASES[name]).items():
            res[f"{name}.{k}"] = v
        print(name, "loss", res[f"{name}.loss"], {k.split(".")[1]: (float(np.nanmax(np.abs(v))), int(np.isnan(v).sum())) for k, v in res.items()
                                                 if k.startswith(name + ".grad")})
    np.savez_compressed(os.path.join(HERE, "train_oov.npz"), **res)
