"""GPU parity (B200): `fdhe` and `dnn` embedders (feat_dh_embedder.py:86-210, dnn_embedder.py:8-112)
through the plugin factory and the C-ABI (`oov_fdhe_embed`) against fixtures generated from the unmodified reference
classes (tests/golden/make_golden_featnet.py) and against the oracle at the tensor-core path's bf16 rounding points."""
import json
import os

import numpy as np
import pytest
import torch

import cases
import parity_util as pu
from oracle import oracle as o

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _golden():
    return np.load(os.path.join(pu.GOLDEN_DIR, "featnet.npz"), allow_pickle=False)


def _embedder(case, inp, tmp_path):
    import gpu_util as G
    import oov_b200
    keys = cases.dhe_keys(case.seed, case.n_hashes)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        os.makedirs("hash_keys", exist_ok=True)
        with open(f"hash_keys/{case.n_hashes}.hashes", "w") as f:
            json.dump([k.hex() for k in keys], f)
        cfg = G.make_config(case, case.kind, user_oov_buckets=4, item_oov_buckets=4, dhe_num_hashes=case.n_hashes,
                            dhe_layer_size=case.layer)
        ds = G.Dataset(4, 4, G.interaction("user_id", inp["user_cols"]), G.interaction("item_id", inp["item_cols"]))
        emb = oov_b200.get_inductive_embedder(cfg, ds, mode=f"test-{case.name}", user_num=4, item_num=4)
    finally:
        os.chdir(cwd)
    with torch.no_grad():
        for side, net in (("user", emb.user_hash_net), ("item", emb.item_hash_net)):
            ws, bs = inp["nets"][side]
            for l, li in enumerate((0, 2, 4, 6)):
                net[li].weight.copy_(G.t(ws[l]))
                net[li].bias.copy_(G.t(bs[l]))
    return emb, keys


@pytest.mark.parametrize("name", list(cases.FEATNET_CASES))
def test_featnet_fp32_path_vs_reference_golden(name, tmp_path):
    import gpu_util as G
    case = cases.FEATNET_CASES[name]
    inp = cases.featnet_inputs(case)
    g = _golden()
    emb, _ = _embedder(case, inp, tmp_path)
    assert type(emb).__name__ == ("FeatDeepHashEmbedder" if case.kind == "fdhe" else "DNNEmbedder")
    assert set(g[f"{name}.state_dict_keys"].tolist()) <= set(emb.state_dict().keys())
    for side in ("user", "item"):
        fm = getattr(emb, f"{side}_feature_mat").cpu().numpy()
        pu.assert_close(fm, g[f"{name}.{side}_feature_mat"], rtol=2e-6, atol=1e-7, what="feature_mat")
    ids = G.t(inp["ids"])
    ids_train = G.t(inp["ids_train"])
    pu.assert_close(emb.embed_item_ids(ids, None).cpu().numpy(), g[f"{name}.item_emb"], rtol=1e-5, atol=1e-6, what="item_emb")
    pu.assert_close(emb.embed_user_ids(ids, None).cpu().numpy(), g[f"{name}.user_emb"], rtol=1e-5, atol=1e-6, what="user_emb")
    emb.set_train()
    before = ids_train.clone()
    pu.assert_close(emb.embed_item_ids(ids_train, None).cpu().numpy(), g[f"{name}.item_emb_train"], rtol=1e-5, atol=1e-6, what="item_emb_train")
    pu.assert_close(emb.embed_user_ids(ids_train, None).cpu().numpy(), g[f"{name}.user_emb_train"], rtol=1e-5, atol=1e-6, what="user_emb_train")
    assert torch.equal(ids_train, before)             # feat_dh_embedder.py:199-201 clones: the caller's ids keep their pad
    emb.set_eval()
    # fused assemble: in-vocab rows gathered, OOV rows embedded
    n_old = case.n_all // 3
    table = G.t(np.arange(n_old * case.D, dtype=np.float32).reshape(n_old, case.D) * 1e-3)
    out = emb.assemble_rows("item", ids, None, n_old, table).cpu().numpy()
    iv = inp["ids"] < n_old
    assert iv.any() and (~iv).any()
    assert (out[iv] == table.cpu().numpy()[inp["ids"][iv]]).all()
    pu.assert_close(out[~iv], g[f"{name}.item_emb"][~iv], rtol=1e-5, atol=1e-6, what="assemble")


@pytest.mark.parametrize("name", list(cases.FEATNET_CASES))
def test_featnet_tensor_core_path(name, tmp_path):
    import gpu_util as G
    from oov_b200 import ops
    case = cases.FEATNET_CASES[name]
    inp = cases.featnet_inputs(case)
    g = _golden()
    emb, keys = _embedder(case, inp, tmp_path)
    emb.compute_path = ops.PATH_TCGEN05
    karr = o.keys_to_array(keys) if case.kind == "fdhe" else None
    for mode, ids_np in (("eval", inp["ids"]), ("train", inp["ids_train"])):
        emb.set_train() if mode == "train" else emb.set_eval()
        ws, bs = inp["nets"]["item"]
        fm = o.featnet_feature_matrix(inp["item_cols"])
        got = emb.embed_item_ids(G.t(ids_np), None).cpu().numpy()
        # oracle at the same rounding points: bf16 weights / features / hidden activations, exact hash inputs
        want = o.fdhe_embed(ids_np, karr, fm, [o.round_bf16(w) for w in ws], bs, training=mode == "train", bf16_points=True)
        pu.assert_close(got, want, rtol=pu.BF16_RTOL, atol=1e-5, what=f"featnet tcgen05 vs oracle(bf16 points) {mode}")
        ref = g[f"{name}.item_emb" + ("_train" if mode == "train" else "")]
        err = np.abs(got - ref).max()
        print(f"[{name} {mode}] tcgen05 vs oracle at bf16 points: {np.abs(got / want - 1).max():.3e}; vs reference fp32 golden: max abs {err:.3e}")
        assert err < 2e-2
    emb.set_eval()
    # bf16 rows straight into a bf16 table (what a bf16 item table holds), in-vocab rows gathered
    n_old = case.n_all // 3
    table = torch.zeros((n_old, case.D), dtype=torch.bfloat16, device=DEV) + 0.25
    out16 = emb.assemble_rows("item", G.t(inp["ids"]), None, n_old, table, out_dtype=torch.bfloat16).float().cpu().numpy()
    iv = inp["ids"] < n_old
    assert (out16[iv] == 0.25).all()
    want = o.fdhe_embed(inp["ids"], karr, o.featnet_feature_matrix(inp["item_cols"]),
                        [o.round_bf16(w) for w in inp["nets"]["item"][0]], inp["nets"]["item"][1], bf16_points=True)
    assert np.abs(out16[~iv] - want[~iv]).max() <= 2 ** -8


def test_featnet_rejects_missing_inputs():
    from oov_b200 import ops
    w = [torch.zeros(64, 12, device=DEV), torch.zeros(64, 64, device=DEV), torch.zeros(64, 64, device=DEV), torch.zeros(8, 64, device=DEV)]
    b = [torch.zeros(64, device=DEV), torch.zeros(64, device=DEV), torch.zeros(64, device=DEV), torch.zeros(8, device=DEV)]
    net = ops.DheNet(w, b, n_feat=4)                      # H = 8 hash inputs + 4 features
    ids = torch.arange(5, device=DEV)
    with pytest.raises(ValueError):
        ops.fdhe_embed(ids, None, net, torch.zeros(5, 4, device=DEV))          # hash inputs but no keys
    with pytest.raises(ValueError):
        ops.fdhe_embed(ids, torch.zeros(8, 16, dtype=torch.uint8, device=DEV), net, torch.zeros(5, 3, device=DEV))   # wrong F
    with pytest.raises(ValueError):
        ops.dhe_embed(ids, torch.zeros(8, 16, dtype=torch.uint8, device=DEV), net)   # the plain-dhe entry refuses feature nets
