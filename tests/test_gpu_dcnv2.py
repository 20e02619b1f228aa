"""DCN-V2 dense tower on the GPU: the product `DCNV2.tower` (tcgen05 linears with folded BatchNorm +
`oov_cross_update`) against the oracle restatement of dcnv2.py:120-144 / 214-250 and the reference-generated golden
(tests/golden/dcnv2_tower.npz, made by tests/golden/make_golden_dcnv2.py from the reference's own modules).

Tolerances: against the oracle evaluated at the kernel's rounding points (bf16 weights, bf16 activations between
layers, fp32 accumulate): 1e-3 relative on the probabilities (north_star's bf16 bound); against the reference's fp32
golden: 5e-3 absolute on the probabilities (observed 5.8e-4 / 8.7e-4) — three chained bf16 cross layers multiply the activations, so the rounding
of every layer compounds; the observed maximum is printed."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as o
from test_oracle_golden import _dcnv2_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "dcnv2_tower.npz")


def _model(structure, x0, cw, cb, layers, pw, pb):
    from oov_b200.model.context import DCNV2
    from oov_b200.inductive.zero_embedder import ZeroEmbedder
    fields = {"stacked": 6, "parallel": 5}[structure]
    D = x0.shape[1] // fields
    cfg = {"embedding_size": D, "structure": structure, "cross_layer_num": 3, "mlp_hidden_size": [L["w"].shape[0] for L in layers],
           "dropout_prob": 0.2, "device": DEV, "mixed": False}
    # users / items 40 in-vocab ids each; ids 40..49 of columns 0 / 1 are out of vocabulary (zero embedder)
    m = DCNV2(cfg, [40, 40] + [50] * (fields - 2), inductive_embedder=ZeroEmbedder(np.zeros((10, 1), np.float32), np.zeros((10, 1), np.float32), 40, 40, D, DEV)).to(DEV).eval()
    sd = m.state_dict()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    for l in range(3):
        sd[f"cross_layer_w.{l}"] = t(cw[l])
        sd[f"bias.{l}"] = t(cb[l]).reshape(-1, 1)
    for l, L in enumerate(layers):       # module indices: Dropout 4l, Linear 4l+1, BatchNorm 4l+2, ReLU 4l+3 (layers.py:60-75)
        sd[f"mlp_layers.mlp_layers.{4 * l + 1}.weight"] = t(L["w"])
        sd[f"mlp_layers.mlp_layers.{4 * l + 1}.bias"] = t(L["b"])
        sd[f"mlp_layers.mlp_layers.{4 * l + 2}.weight"] = t(L["bn_gamma"])
        sd[f"mlp_layers.mlp_layers.{4 * l + 2}.bias"] = t(L["bn_beta"])
        sd[f"mlp_layers.mlp_layers.{4 * l + 2}.running_mean"] = t(L["bn_mean"])
        sd[f"mlp_layers.mlp_layers.{4 * l + 2}.running_var"] = t(L["bn_var"])
    sd["predict_layer.weight"] = t(pw).reshape(1, -1)
    sd["predict_layer.bias"] = t(pb).reshape(1)
    m.load_state_dict(sd)
    m.pack_tower()
    return m


@pytest.mark.parametrize("structure", ["stacked", "parallel"])
def test_dcnv2_tower_vs_oracle_and_reference_golden(structure):
    g = np.load(GOLD)
    x0, cw, cb, layers, pw, pb = _dcnv2_case(g, structure)
    m = _model(structure, x0, cw, cb, layers, pw, pb)
    x16 = torch.from_numpy(x0).to(DEV).to(torch.bfloat16)
    got = m.tower(x16).float().cpu().numpy()
    assert got.shape == (x0.shape[0],)
    # oracle at the kernel's rounding points: bf16 inputs and weights (BatchNorm folded BEFORE rounding, like pack_tower)
    r = o.round_bf16
    folded = [o.fold_bn(L) for L in layers]
    ident = [dict(w=r(w), b=b, bn_mean=np.zeros_like(b), bn_var=np.ones_like(b), bn_gamma=np.ones_like(b), bn_beta=np.zeros_like(b), bn_eps=0.0)
             for w, b in folded]
    want16 = o.dcnv2_forward(r(x0), [r(w) for w in cw], cb, ident, r(pw), pb, structure, bf16_points=True)
    err16 = np.abs(got - want16) / np.maximum(np.abs(want16), 1e-6)
    want32 = g[structure + ".out"]
    err32 = np.abs(got - want32)
    print(f"[dcnv2 {structure}] vs oracle at bf16 points: max rel {err16.max():.3e}; vs reference fp32 golden: max abs {err32.max():.3e}")
    assert err16.max() <= 1e-3, err16.max()
    assert err32.max() <= 5e-3, err32.max()


def test_dcnv2_cross_update_and_relu_epilogue():
    """`oov_cross_update` is x0 * t + xl in fp32 rounded once to bf16 (bit-exact against torch); the ReLU epilogue of
    `oov_tc_linear` equals relu of the plain epilogue."""
    from oov_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(5)
    for n, d in ((1, 8), (777, 40), (4096, 416)):
        x0, t, xl = (torch.randn((n, d), generator=gen).to(DEV).to(torch.bfloat16) for _ in range(3))
        got = ops.cross_update(x0, t, xl)
        want = (x0.float() * t.float() + xl.float()).to(torch.bfloat16)
        assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    A = torch.randn((1000, 96), generator=gen).to(DEV).to(torch.bfloat16)
    W = torch.randn((416, 96), generator=gen).to(DEV).to(torch.bfloat16)
    b = torch.randn((416,), generator=gen).to(DEV)
    for dt in (torch.float32, torch.bfloat16):
        plain = ops.tc_linear(A, W, b, act="none", out_dtype=torch.float32)
        relu = ops.tc_linear(A, W, b, act="relu", out_dtype=dt)
        assert torch.equal(relu, torch.relu(plain).to(dt))
    ref = A.float() @ W.float().T + b
    assert torch.allclose(plain, ref, rtol=1e-4, atol=1e-3)


def test_tc_linear_fast_path_widths_not_multiple_of_64():
    """The staged-store epilogue now takes any N % 8 == 0 in (64, 1024]: compare with fp32 maths for ragged widths."""
    from oov_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(10)
    for M, N, K in ((257, 72, 64), (1000, 416, 416), (333, 264, 128), (4096, 1016, 96)):
        A = torch.randn((M, K), generator=gen).to(DEV).to(torch.bfloat16)
        W = torch.randn((N, K), generator=gen).to(DEV).to(torch.bfloat16)
        b = torch.randn((N,), generator=gen).to(DEV)
        for act, fn in (("none", lambda x: x), ("relu", torch.relu), ("gelu", torch.nn.functional.gelu)):
            got = ops.tc_linear(A, W, b, act=act, out_dtype=torch.bfloat16).float()
            want = fn(A.float() @ W.float().T + b)
            assert torch.allclose(got, want, rtol=2.0 ** -7, atol=2e-2), (M, N, K, act, (got - want).abs().max().item())


def test_graphed_ranker_replays_the_forward():
    """GraphedRanker: the captured forward equals the eager one bit for bit, for full and short batches, from host or
    device token tensors."""
    import oov_b200
    g = np.load(GOLD)
    m = _model("stacked", *_dcnv2_case(g, "stacked"))
    gen = torch.Generator(device="cpu").manual_seed(4)
    gr = oov_b200.GraphedRanker(m, 512, 6)
    assert gr.launches_per_replay >= 6
    for n in (512, 100, 1):
        tokens = torch.randint(0, 50, (n, 6), generator=gen)
        want = m(tokens.to(DEV))
        got = gr(tokens.pin_memory()).clone()
        got_dev = gr(tokens.to(DEV)).clone()
        torch.cuda.synchronize()
        assert torch.equal(got, want) and torch.equal(got_dev, want)
    with pytest.raises(ValueError):
        gr(torch.zeros((513, 6), dtype=torch.int64))


def test_dcnv2_forward_through_token_gather_and_oov_overwrite():
    """End to end: token ids (some out of vocabulary) -> embed_token_fields -> tower, against the oracle tower applied to
    the same gathered embeddings."""
    from oov_b200.model.context import DCNV2
    g = np.load(GOLD)
    x0, cw, cb, layers, pw, pb = _dcnv2_case(g, "stacked")
    m = _model("stacked", x0, cw, cb, layers, pw, pb)
    gen = torch.Generator(device="cpu").manual_seed(3)
    tokens = torch.randint(0, 50, (300, 6), generator=gen).to(DEV)
    emb = m.embed_token_fields(tokens, out_dtype=torch.bfloat16)
    oov = tokens[:, 0] >= 40
    assert oov.any() and (emb[oov, 0].float() == 0).all()           # OOV user cells were overwritten by the embedder
    x = emb.reshape(300, -1)
    got = m(tokens).cpu().numpy()
    r = o.round_bf16
    folded = [o.fold_bn(L) for L in layers]
    ident = [dict(w=r(w), b=b, bn_mean=np.zeros_like(b), bn_var=np.ones_like(b), bn_gamma=np.ones_like(b), bn_beta=np.zeros_like(b), bn_eps=0.0)
             for w, b in folded]
    want = o.dcnv2_forward(x.float().cpu().numpy(), [r(w) for w in cw], cb, ident, r(pw), pb, "stacked", bf16_points=True)
    assert np.abs(got - want).max() <= 1e-3 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("name", ["default", "wide"])
def test_widedeep_head_vs_oracle_and_reference_golden(name):
    """`WideDeep.deep` (tcgen05 linears with ReLU epilogue, widths 32 / 16 / 8 and 256 / 128, K = 260 padded to 264) + the
    wide term against the oracle at the kernel's rounding points (bf16 input, weights and hidden activations) and the
    reference's fp32 golden (tests/golden/widedeep_head.npz)."""
    from oov_b200.model.context import WideDeep
    from oov_b200.inductive.zero_embedder import ZeroEmbedder
    from test_oracle_golden import _widedeep_case
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "widedeep_head.npz"))
    emb, fm, ws, bs, pw, pb = _widedeep_case(g, name)
    Bn, fields, D = emb.shape
    cfg = {"embedding_size": D, "mlp_hidden_size": [w.shape[0] for w in ws], "dropout_prob": 0.1, "device": DEV}
    z = lambda d: ZeroEmbedder(np.zeros((10, 1), np.float32), np.zeros((10, 1), np.float32), 40, 40, d, DEV)
    m = WideDeep(cfg, [40, 40] + [50] * (fields - 2), inductive_embedder=z(D), first_order_embedder=z(1)).to(DEV).eval()
    sd = m.state_dict()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    for l, (w, b) in enumerate(zip(ws, bs)):     # module indices without BatchNorm: Dropout 3l, Linear 3l+1, ReLU 3l+2 (layers.py:60-75)
        sd[f"mlp_layers.mlp_layers.{3 * l + 1}.weight"] = t(w)
        sd[f"mlp_layers.mlp_layers.{3 * l + 1}.bias"] = t(b)
    sd["deep_predict_layer.weight"] = t(pw).reshape(1, -1)
    sd["deep_predict_layer.bias"] = t(pb).reshape(1)
    m.load_state_dict(sd)
    m.pack_tower()
    x16 = torch.from_numpy(emb.reshape(Bn, -1)).to(DEV).to(torch.bfloat16)
    logits = (torch.from_numpy(fm.reshape(-1)).to(DEV) + m.deep(x16)).cpu().numpy()
    r = o.round_bf16
    want16 = o.widedeep_forward(r(emb), fm, [r(w) for w in ws], bs, r(pw), pb, bf16_points=True)
    assert np.abs(logits - want16).max() <= 1e-3 * max(1.0, np.abs(want16).max()), np.abs(logits - want16).max()
    err32 = np.abs(o.sigmoid(logits) - g[name + ".prob"]).max()
    print(f"[widedeep {name}] vs oracle at bf16 points: {np.abs(logits - want16).max():.3e}; probabilities vs reference fp32 golden: {err32:.3e}")
    assert err32 <= 5e-3
    # whole forward through the token gather and both OOV embedders: finite, wide + deep, sigmoid in (0, 1)
    tokens = torch.randint(0, 50, (257, fields), generator=torch.Generator().manual_seed(1)).to(DEV)
    p = m.predict(tokens)
    emb_t = m.embed_token_fields(tokens, out_dtype=torch.bfloat16)
    want_p = torch.sigmoid(m.first_order_linear(tokens).reshape(-1) + m.deep(emb_t.reshape(257, -1)))
    assert p.shape == (257,) and torch.equal(p, want_p) and bool(((p > 0) & (p < 1)).all())
