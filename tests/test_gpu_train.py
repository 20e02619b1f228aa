"""Training-mode OOV path on the GPU: `calculate_loss` + backward of the product models (assemble kernels
under autograd, `oov_scatter_add_rows` / `oov_lsh_embed_backward`) against the loss and the table gradients of the
UNMODIFIED reference models in `set_oov_train()` mode (tests/golden/train_oov.npz, made by make_golden_train.py), plus
kernel-level checks against the oracle at larger sizes.

Tolerance: fp32 rtol 1e-4 / atol 1e-7 on the gradients (sums of up to a few hundred fp32 terms in atomics order), the loss
at rtol 1e-5; NaN patterns (all-zero LSH hashes) must coincide."""
import os

import numpy as np
import pytest
import torch

import cases
import parity_util as pu
from oracle import oracle as o

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(pu.GOLDEN_DIR, "train_oov.npz")


@pytest.mark.parametrize("name", cases.TRAIN_CASES)
def test_calculate_loss_and_table_gradients_vs_reference(name):
    import gpu_util as G
    import oov_b200
    g = np.load(GOLD)
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    cfg, emb, model = G.build_retrieval(case, inp)
    model.train()
    model.set_oov_train()
    b = cases.train_batch(case)
    inter = oov_b200.Interaction({"user_id": G.t(b["users"]), "item_id": G.t(b["pos"]), "neg_item_id": G.t(b["neg"])})
    before = inter["user_id"].clone()
    loss = model.calculate_loss(inter)
    loss.backward()
    assert torch.equal(inter["user_id"], before)                     # bpr.py:64 embeds a masked COPY: the batch keeps its pads
    want_loss = float(g[f"{name}.loss"])
    if np.isnan(want_loss):
        assert np.isnan(loss.item())
    else:
        assert abs(loss.item() - want_loss) <= 1e-5 * max(1.0, abs(want_loss)), (loss.item(), want_loss)
    for nm in ("user_embedding", "item_embedding", "user_oov_buckets", "item_oov_buckets"):
        gr = getattr(model, nm).weight.grad
        got = np.zeros_like(g[f"{name}.grad_{nm}"]) if gr is None else gr.cpu().numpy()
        want = g[f"{name}.grad_{nm}"]
        assert (np.isnan(got) == np.isnan(want)).all(), f"{name} {nm}: NaN pattern differs"
        ok = ~np.isnan(want)
        scale = np.abs(want[ok]).max() if ok.any() else 0.0
        assert np.abs(got[ok] - want[ok]).max(initial=0.0) <= 1e-4 * scale + 1e-7, (name, nm, np.abs(got[ok] - want[ok]).max(), scale)
    model.set_oov_eval()
    # eval-mode calls stay outside autograd
    assert not model.get_user_embedding(G.t(b["users"][:4] % case.n_old_users)).requires_grad


def test_oov_freeze_embedding_stops_table_gradients():
    import gpu_util as G
    import oov_b200
    case = cases.CASES["bpr_slsh_odd"]
    inp = cases.retrieval_inputs(case)
    cfg, emb, model = G.build_retrieval(case, inp)
    model.oov_freeze_embedding = True
    model.train()
    model.set_oov_train()                                            # freezes user / item tables (bpr.py:86-92)
    b = cases.train_batch(case)
    inter = oov_b200.Interaction({"user_id": G.t(b["users"]), "item_id": G.t(b["pos"]), "neg_item_id": G.t(b["neg"])})
    model.calculate_loss(inter).backward()
    assert model.user_embedding.weight.grad is None and model.item_embedding.weight.grad is None
    assert model.user_oov_buckets.weight.grad is not None and float(model.user_oov_buckets.weight.grad.abs().sum()) > 0
    model.set_oov_eval()
    assert model.user_embedding.weight.requires_grad


@pytest.mark.parametrize("n,B,D", [(5000, 1000, 64), (777, 70, 24), (300, 1000, 100)])
def test_lsh_backward_kernel_vs_oracle(n, B, D):
    from oov_b200 import ops
    rs = np.random.RandomState(n + B)
    H = (rs.rand(n, B) < 0.5).astype(np.uint8)
    H[5] = 0                                                         # an all-zero hash ...
    ids = rs.randint(0, 2000, size=n).astype(np.int64)
    n_old = 700
    ids[5] = 0                                                       # ... on an in-vocab row: must not poison dW
    gnp = rs.randn(n, D).astype(np.float32)
    words = np.packbits(np.pad(H, ((0, 0), (0, (-B) % 32))), axis=1, bitorder="little").view(np.uint32).astype(np.int64)
    bits = torch.from_numpy(words.astype(np.uint32).view(np.int32).reshape(n, -1)).to(DEV)
    if B % 32:
        bits[:, -1] |= torch.tensor(-(1 << (B % 32)), dtype=torch.int32, device=DEV)    # garbage above bit B must be ignored
    dW = torch.zeros((B, D), dtype=torch.float32, device=DEV)
    ops.lsh_embed_backward(bits, torch.from_numpy(gnp).to(DEV), torch.from_numpy(ids).to(DEV), n_old, dW)
    oov = ids >= n_old
    want = o.lsh_embed_backward(H[oov], gnp[oov])
    pu.assert_close(dW.cpu().numpy(), want, rtol=1e-4, atol=1e-5, what="lsh backward")
    # additive (the entry point ADDS into dW) and linear in g
    ops.lsh_embed_backward(bits, torch.from_numpy(2 * gnp).to(DEV), torch.from_numpy(ids).to(DEV), n_old, dW)
    pu.assert_close(dW.cpu().numpy(), 3 * want, rtol=1e-4, atol=1e-5, what="lsh backward additive")
    # an all-zero hash on an OOV row poisons every bucket with NaN (autograd's 0 * inf)
    ids[5] = n_old + 1
    dW.zero_()
    ops.lsh_embed_backward(bits, torch.from_numpy(gnp).to(DEV), torch.from_numpy(ids).to(DEV), n_old, dW)
    assert bool(torch.isnan(dW).all())


def test_scatter_add_rows_vs_oracle():
    from oov_b200 import ops
    rs = np.random.RandomState(3)
    n, rows, D = 20000, 300, 48
    idx = rs.randint(-50, 400, size=n).astype(np.int64)
    gnp = rs.randn(n, D).astype(np.float32)
    dt = torch.zeros((rows, D), dtype=torch.float32, device=DEV)
    ops.scatter_add_rows(torch.from_numpy(gnp).to(DEV), torch.from_numpy(idx).to(DEV), dt)
    pu.assert_close(dt.cpu().numpy(), o.scatter_add_rows(gnp, idx, rows), rtol=1e-4, atol=1e-4, what="scatter add")
    dt.zero_()
    ops.scatter_add_rows(torch.from_numpy(gnp).to(DEV), torch.from_numpy(idx).to(DEV), dt, idx_offset=-100)
    pu.assert_close(dt.cpu().numpy(), o.scatter_add_rows(gnp, idx - 100, rows), rtol=1e-4, atol=1e-4, what="scatter add offset")


@pytest.mark.parametrize("name", list(cases.HASHNET_TRAIN_CASES))
def test_hash_net_training_vs_reference(name, tmp_path):
    """BPR + dhe / fdhe / dnn in `set_oov_train()` mode: loss, table gradients and the gradients of all 16 hash-net
    parameters against the unmodified reference under autograd (fp32 path, rtol 2e-4 of each tensor's scale)."""
    import json
    import gpu_util as G
    import oov_b200
    g = np.load(GOLD)
    case = cases.HASHNET_TRAIN_CASES[name]
    inp = cases.hashnet_train_inputs(case)
    n_old = inp["n_old"]
    keys = cases.dhe_keys(case.seed, case.n_hashes)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        os.makedirs("hash_keys", exist_ok=True)
        with open(f"hash_keys/{case.n_hashes}.hashes", "w") as f:
            json.dump([k.hex() for k in keys], f)
        cfg = G.make_config(case, case.kind, user_oov_buckets=4, item_oov_buckets=4, dhe_num_hashes=case.n_hashes, dhe_layer_size=case.layer)
        ds = G.Dataset(n_old, n_old, G.interaction("user_id", inp["user_cols"]), G.interaction("item_id", inp["item_cols"]))
        emb = oov_b200.get_inductive_embedder(cfg, ds, mode=f"test-train-{name}", user_num=n_old, item_num=n_old)
    finally:
        os.chdir(cwd)
    model = oov_b200.BPR(cfg, ds, inductive_mapper=None, inductive_embedder=emb).to(DEV)
    with torch.no_grad():
        model.user_embedding.weight.copy_(G.t(inp["user_table"]))
        model.item_embedding.weight.copy_(G.t(inp["item_table"]))
        for side, net in (("user", emb.user_hash_net), ("item", emb.item_hash_net)):
            ws, bs = inp["nets"][side]
            for l, li in enumerate((0, 2, 4, 6)):
                net[li].weight.copy_(G.t(ws[l]))
                net[li].bias.copy_(G.t(bs[l]))
    model.train()
    model.set_oov_train()
    b = inp["batch"]
    inter = oov_b200.Interaction({"user_id": G.t(b["users"]), "item_id": G.t(b["pos"]), "neg_item_id": G.t(b["neg"])})
    loss = model.calculate_loss(inter)
    loss.backward()
    want_loss = float(g[f"{name}.loss"])
    assert abs(loss.item() - want_loss) <= 1e-5 * max(1.0, abs(want_loss)), (loss.item(), want_loss)
    got = {"user_embedding": model.user_embedding.weight.grad, "item_embedding": model.item_embedding.weight.grad}
    for side, net in (("user", emb.user_hash_net), ("item", emb.item_hash_net)):
        for li in (0, 2, 4, 6):
            got[f"{side}_hash_net.{li}.weight"] = net[li].weight.grad
            got[f"{side}_hash_net.{li}.bias"] = net[li].bias.grad
    worst = 0.0
    for nm, gr in got.items():
        want = g[f"{name}.grad_{nm}"]
        assert gr is not None, nm
        have = cases.grad_slice(gr.cpu().numpy()) if nm.endswith("weight") and "hash_net" in nm else gr.cpu().numpy()
        assert have.shape == want.shape, (nm, have.shape, want.shape)
        scale = np.abs(want).max()
        err = np.abs(have - want).max()
        worst = max(worst, err / max(scale, 1e-30))
        assert err <= 2e-4 * scale + 1e-9, (name, nm, err, scale)
    print(f"[{name}] loss {loss.item():.6f}; worst gradient error / tensor scale {worst:.2e}")
