"""Golden fixtures of the `fdhe` and `dnn` embedders: the UNMODIFIED reference classes (FeatDeepHashEmbedder,
DNNEmbedder) built through the reference's own factory on seeded inputs.  Authoring container only (needs /root/reference):
    python tests/golden/make_golden_featnet.py   ->   tests/golden/featnet.npz
csiphash (absent here) is served by oracle/siphash24.c through oracle/refshim.py, like for the dhe fixtures."""
import json
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
from make_golden import base_config, interaction, ns, T  # noqa: E402


def run(case: cases.FeatNetCase) -> dict:
    inp = cases.featnet_inputs(case)
    keys = cases.dhe_keys(case.seed, case.n_hashes)
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                                   # ./hash_keys is CWD-relative (feat_dh_embedder.py:52,131-136)
        try:
            os.makedirs("hash_keys")
            with open(f"hash_keys/{case.n_hashes}.hashes", "w") as f:
                json.dump([k.hex() for k in keys], f)
            cfg = base_config(case, case.kind, user_oov_buckets=4, item_oov_buckets=4, dhe_num_hashes=case.n_hashes,
                              dhe_layer_size=case.layer)
            ds = refshim.RefDataset(4, 4, interaction("user_id", inp["user_cols"]), interaction("item_id", inp["item_cols"]))
            emb = ns.get_inductive.get_inductive_embedder(cfg, ds, mode=f"golden-{case.name}", user_num=4, item_num=4)
            assert type(emb).__name__ == ("FeatDeepHashEmbedder" if case.kind == "fdhe" else "DNNEmbedder")
            out["state_dict_keys"] = np.array(sorted(emb.state_dict().keys()))
            out["user_feature_mat"] = emb.user_feature_mat.numpy()
            out["item_feature_mat"] = emb.item_feature_mat.numpy()
            with torch.no_grad():
                for side, net in (("user", emb.user_hash_net), ("item", emb.item_hash_net)):
                    ws, bs = inp["nets"][side]
                    for l, li in enumerate((0, 2, 4, 6)):
                        assert tuple(net[li].weight.shape) == ws[l].shape
                        net[li].weight.copy_(T(ws[l]))
                        net[li].bias.copy_(T(bs[l]))
                emb.set_eval()
                out["item_emb"] = emb.embed_item_ids(T(inp["ids"]), None).numpy()
                out["user_emb"] = emb.embed_user_ids(T(inp["ids"]), None).numpy()
                emb.set_train()
                out["item_emb_train"] = emb.embed_item_ids(T(inp["ids_train"].copy()), None).numpy()
                out["user_emb_train"] = emb.embed_user_ids(T(inp["ids_train"].copy()), None).numpy()
                emb.set_eval()
        finally:
            os.chdir(cwd)
    return out


if __name__ == "__main__":
    res = {}
    for name, case in cases.FEATNET_CASES.items():
        for k, v in run(case).items():
            res[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(HERE, "featnet.npz"), **res)
    print("wrote featnet.npz", {k: v.shape for k, v in res.items() if k.endswith("item_emb")})
