"""GPU parity (B200): the tcgen05/TMA/TMEM linear layer and the tensor-core DHE path.

bf16 operands, fp32 accumulation: compared with a plain fp32 torch reference of the same op on the
same bf16-rounded inputs (tolerance rtol 1e-3, the north star's bf16 contract; the observed error is
accumulation-order only), and with the oracle evaluated at the same bf16 rounding points."""
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


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 512, 384), (300, 64, 512), (128, 16, 96), (77, 512, 512),
                                   (4097, 512, 128), (1, 64, 64)])
@pytest.mark.parametrize("act", ["none", "gelu", "sigmoid"])
def test_tc_linear_matches_fp32_reference(M, N, K, act):
    from oov_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    got = ops.tc_linear(A, W, bias, act=act)
    ref = A.float() @ W.float().T + bias
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
    elif act == "sigmoid":
        ref = torch.sigmoid(ref)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    assert torch.allclose(got, ref, rtol=1e-3, atol=1e-4), f"max abs err {err}"
    # bf16 output = rounding of the fp32 result (1 ulp slack for results that land on a rounding boundary)
    got16 = ops.tc_linear(A, W, bias, act=act, out_dtype=torch.bfloat16).float()
    assert torch.allclose(got16, ref, rtol=2 ** -7, atol=1e-3)


def test_tc_linear_exact_small_integers():
    """Products and sums of small integers are exact in bf16 x bf16 -> fp32: catches any operand-layout /
    descriptor / swizzle error bit-for-bit (each output depends on a distinct row/column pattern)."""
    from oov_b200 import ops
    M, N, K = 256, 512, 192
    A = ((torch.arange(M).view(-1, 1) * 3 + torch.arange(K).view(1, -1) * 5) % 7 - 3).float()
    W = ((torch.arange(N).view(-1, 1) * 11 + torch.arange(K).view(1, -1) * 13) % 5 - 2).float()
    got = ops.tc_linear(A.to(torch.bfloat16).to(DEV), W.to(torch.bfloat16).to(DEV), None, act="none")
    assert torch.equal(got.cpu(), A @ W.T)


def _dhe(case, tmp_path, path):
    import gpu_util as G
    import oov_b200
    from oov_b200 import ops
    keys = cases.dhe_keys(case.seed, case.n_hashes)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        os.makedirs("hash_keys", exist_ok=True)
        with open(f"hash_keys/{case.n_hashes}.hashes", "w") as f:
            json.dump([k.hex() for k in keys], f)
        fu = oov_b200.Interaction({"user_id": torch.arange(8), "f0": torch.ones(8, 2)})
        fi = oov_b200.Interaction({"item_id": torch.arange(8), "f0": torch.ones(8, 2)})
        cfg = G.make_config(case, "dhe", user_oov_buckets=4, item_oov_buckets=4, dhe_num_hashes=case.n_hashes)
        emb = oov_b200.get_inductive_embedder(cfg, G.Dataset(4, 4, fu, fi), mode=f"tc-{case.name}", user_num=4, item_num=4)
    finally:
        os.chdir(cwd)
    ws, bs = cases.dhe_weights(case)
    with torch.no_grad():
        for l, li in enumerate((0, 2, 4, 6)):
            emb.item_hash_net[li].weight.copy_(G.t(ws[l]))
            emb.item_hash_net[li].bias.copy_(G.t(bs[l]))
    emb.compute_path = path
    return emb, keys, ws, bs


@pytest.mark.parametrize("name", list(cases.DHE_CASES))
def test_dhe_tensor_core_path(name, tmp_path):
    import gpu_util as G
    from oov_b200 import ops
    case = cases.DHE_CASES[name]
    g = pu.load_golden(name)
    emb, keys, ws, bs = _dhe(case, tmp_path, ops.PATH_TCGEN05)
    ids_np = cases.dhe_ids(case)
    ids = G.t(ids_np)
    got = emb.embed_item_ids(ids, None).cpu().numpy()
    # oracle at the same rounding points: bf16 weights, bf16 hidden activations, exact 24-bit inputs, fp32 accumulate
    h = o.dhe_hashes(ids_np, o.keys_to_array(keys))
    want = o.dhe_mlp(h, [o.round_bf16(w) for w in ws], bs, bf16_points=True)
    if case.w1_scale == 1.0:
        sat = np.abs(g["item_logits"]) > 1e4            # far inside saturation: bf16 weight rounding cannot flip the sign
        assert sat.mean() > 0.9
        assert (got[sat] == g["item_emb"][sat]).all()
    else:
        pu.assert_close(got, want, rtol=pu.BF16_RTOL, atol=1e-5, what="dhe tcgen05 vs oracle(bf16 points)")
        # and against the fp32 reference itself, at the looser tolerance bf16 weights allow (reported, not the contract)
        err = np.abs(got - g["item_emb"]).max()
        print(f"[{name}] tcgen05 DHE vs fp32 reference: max abs err {err:.3e}")
        assert err < 2e-2
    # bf16 output (what bf16 item tables hold) = the fp32 result rounded
    n_old = 50
    table = torch.zeros((n_old, case.D), dtype=torch.bfloat16, device=DEV) + 0.25
    out16 = emb.assemble_rows("item", ids, None, n_old, table, out_dtype=torch.bfloat16).float().cpu().numpy()
    iv = ids_np < n_old
    assert (out16[iv] == 0.25).all()
    if case.w1_scale != 1.0:
        assert np.abs(out16[~iv] - want[~iv]).max() <= 2 ** -8


def test_dhe_tc_large_matches_simt(tmp_path):
    """Larger n (several tiles per CTA, ragged last tile): tensor-core path vs the fp32 CUDA-core path."""
    import gpu_util as G
    from oov_b200 import ops
    case = cases.DHE_CASES["dhe_scaled"]
    emb, keys, ws, bs = _dhe(case, tmp_path, ops.PATH_TCGEN05)
    ids = torch.arange(1000, 1000 + 40_001, device=DEV)
    a = emb.embed_item_ids(ids, None)
    emb.compute_path = ops.PATH_SIMT_FP32
    b = emb.embed_item_ids(ids, None)
    torch.cuda.synchronize()
    assert (a - b).abs().max().item() < 5e-3
    assert (a - b).abs().mean().item() < 5e-4
