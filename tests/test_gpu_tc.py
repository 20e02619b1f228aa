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


# ------------------------------------------------------------------------------------------------ fused tcgen05 score + top-k
def _ref_scores(users, items, off, hist, seg, mask_pad=True):
    s = (users.float() @ items.float().T).cpu().numpy()
    n = s.shape[1]
    gid = np.arange(n) + off
    if mask_pad:
        s[:, gid == 0] = -np.inf
    s[:, (gid < seg[0]) | (gid >= seg[1])] = -np.inf
    if hist is not None:
        hu, hi = hist
        loc = hi - off
        ok = (loc >= 0) & (loc < n)
        s[hu[ok], loc[ok]] = -np.inf
    return s


@pytest.mark.parametrize("Q,N,D,k", [(1, 50, 64, 5), (100, 257, 64, 20), (256, 5000, 64, 20), (300, 5000, 16, 24),
                                     (1024, 70_001, 64, 20), (129, 1024, 32, 1), (64, 200_003, 64, 10)])
def test_tc_score_topk_vs_fp32_reference(Q, N, D, k):
    from oov_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(Q + N + D + k)
    users = torch.randn(Q, D, generator=g).to(torch.bfloat16).to(DEV)
    items = torch.randn(N, D, generator=g).to(torch.bfloat16).to(DEV)
    if N > 40:
        items[10:20] = items[30:40]                        # exact score ties -> (score desc, id asc) must hold
    hu = torch.randint(0, Q, (min(5 * Q, 3000),), generator=g)
    hi = torch.randint(0, N, (hu.numel(),), generator=g) + 1000
    off = 1000
    hist = ops.pairs_to_csr(hu.to(DEV), hi.to(DEV), Q)
    for seg in ((0, 1 << 62), (off + N // 3, off + 2 * N // 3)):
        s_tc, i_tc = ops.fullsort_topk(users, items, k, item_id_offset=off, hist=hist, seg=seg, path=ops.PATH_TCGEN05)
        s_si, i_si = ops.fullsort_topk(users, items, k, item_id_offset=off, hist=hist, seg=seg, path=ops.PATH_SIMT_FP32)
        torch.cuda.synchronize()
        ref = _ref_scores(users, items, off, (hu.numpy(), hi.numpy()), seg)
        scale = float(np.abs(ref[np.isfinite(ref)]).max())
        for nm, s_, i_ in (("tcgen05", s_tc, i_tc), ("simt", s_si, i_si)):
            idx = i_.cpu().numpy() - off
            kk = min(k, N)
            ok, msg = o.topk_sets_match(ref, idx[:, :kk], kk, rtol=1e-5, atol=1e-5 * scale)
            assert ok, f"{nm} seg={seg}: {msg}"
            picked = np.take_along_axis(ref, idx[:, :kk], axis=1)
            pu.assert_close(s_.cpu().numpy()[:, :kk], picked, rtol=1e-5, atol=1e-5 * scale, what=f"{nm} scores")
            if kk < k:
                assert (idx[:, kk:] == -1 - off).all()
            key = o.order_key(s_.cpu().numpy())
            assert (key[:, :-1] >= key[:, 1:]).all(), nm
            tie = key[:, :-1] == key[:, 1:]
            assert (i_.cpu().numpy()[:, :-1][tie] < i_.cpu().numpy()[:, 1:][tie]).all(), nm


def test_tc_score_prepass_history_holds_the_top_items():
    """Adversarial case for the sampled pre-pass threshold (tc_score.cu): every user's history is exactly the user's
    40 best items, so most sampled tile maxima belong to masked items; plus NaN items and a pad row in a sampled tile."""
    from oov_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(7)
    Q, N, D, k = 96, 100_000, 64, 20
    users = torch.randn(Q, D, generator=g).to(torch.bfloat16).to(DEV)
    items = torch.randn(N, D, generator=g).to(torch.bfloat16).to(DEV)
    items[[77, 50_000]] = float("nan")
    full = users.float() @ items.float().T
    top = torch.topk(torch.nan_to_num(full, nan=-1e30), 40, dim=1).indices          # [Q, 40]
    hu = torch.arange(Q, device=DEV).repeat_interleave(40)
    hi = top.reshape(-1)
    hist = ops.pairs_to_csr(hu, hi, Q)
    ref = full.clone()
    ref[:, 0] = -float("inf")
    ref[hu, hi] = -float("inf")
    for seg in ((0, 1 << 62), (1000, 90_000)):
        r = ref.clone()
        r[:, :seg[0]] = -float("inf")
        r[:, min(seg[1], N):] = -float("inf")
        r = torch.nan_to_num(r, nan=float("inf"), posinf=float("inf"), neginf=-float("inf"))    # NaN ranks first
        s_tc, i_tc = ops.fullsort_topk(users, items, k, hist=hist, seg=seg, path=ops.PATH_TCGEN05)
        s_si, i_si = ops.fullsort_topk(users, items, k, hist=hist, seg=seg, path=ops.PATH_SIMT_FP32)
        ok, msg = o.topk_sets_match(r.cpu().numpy(), i_tc.cpu().numpy(), k, rtol=1e-5, atol=1e-4)
        assert ok, f"seg={seg}: {msg}"
        want = [77, 50_000] if seg[0] == 0 else [50_000]
        assert torch.equal(i_tc[:, :len(want)].cpu(), torch.tensor(want).expand(Q, len(want)))   # NaN first, id ascending
        # SIMT fp32 accumulates in a different order: index sets agree up to near-ties, checked against ref above
        ok, msg = o.topk_sets_match(r.cpu().numpy(), i_si.cpu().numpy(), k, rtol=1e-5, atol=1e-4)
        assert ok, f"simt seg={seg}: {msg}"


@pytest.fixture
def main2_env():
    old = os.environ.get("OOV_SCORE_MAIN2")
    yield lambda v: os.environ.__setitem__("OOV_SCORE_MAIN2", str(v))
    if old is None:
        os.environ.pop("OOV_SCORE_MAIN2", None)
    else:
        os.environ["OOV_SCORE_MAIN2"] = old


@pytest.mark.parametrize("Q,N,k", [(96, 100_000, 20), (700, 70_001, 24), (1024, 300_000, 20), (513, 40_000, 1)])
def test_tc_score_column_split_main_pass_equals_thread_per_user_main_pass(Q, N, k, main2_env):
    """The two main passes behind the sampled threshold (tc_score.cu: MODE 0, thread = user; tc_score_main2_kernel, scan
    warps + collector warps, round-robin tiles) give the same lists bit for bit — scores, ids and order — on inputs with
    exact score ties, NaN rows, a pad row, histories made of each user's best items (so sampled maxima are masked items and
    the collectors' history test is exercised), item-id offsets and kept segments that cut through tiles; and both match
    the fp32 reference.  OOV_SCORE_MAIN2 = 2 forces the column-split kernel at sizes where the automatic rule would not
    pick it (it is read on every call)."""
    from oov_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(Q + N + k)
    D = 64
    users = torch.randn(Q, D, generator=g).to(torch.bfloat16).to(DEV)
    items = torch.randn(N, D, generator=g).to(torch.bfloat16).to(DEV)
    items[10:20] = items[30:40]                            # exact ties
    items[N // 2: N // 2 + 300] = items[5:305]             # more ties, in another CTA's tiles
    items[[77, N - 3]] = float("nan")
    off = 1000
    full = users.float() @ items.float().T
    top = torch.topk(torch.nan_to_num(full, nan=-1e30), 30, dim=1).indices
    hu = torch.cat([torch.arange(Q, device=DEV).repeat_interleave(30), torch.randint(0, Q, (4 * Q,), generator=g).to(DEV)])
    hi = torch.cat([top.reshape(-1), torch.randint(0, N, (4 * Q,), generator=g).to(DEV)]) + off
    hist = ops.pairs_to_csr(hu, hi, Q)
    for seg in ((0, 1 << 62), (off + N // 5 + 17, off + N - N // 7)):
        main2_env(0)
        s0, i0 = ops.fullsort_topk(users, items, k, item_id_offset=off, hist=hist, seg=seg, path=ops.PATH_TCGEN05)
        main2_env(2)
        s2, i2 = ops.fullsort_topk(users, items, k, item_id_offset=off, hist=hist, seg=seg, path=ops.PATH_TCGEN05)
        torch.cuda.synchronize()
        assert torch.equal(i0, i2), f"seg={seg}: {(i0 != i2).sum().item()} ids differ"
        assert torch.equal(s0.view(torch.int32), s2.view(torch.int32)), f"seg={seg}: scores differ"
        ref = _ref_scores(users, items, off, (hu.cpu().numpy(), (hi - 0).cpu().numpy()), seg)
        ref = np.nan_to_num(ref, nan=np.inf, posinf=np.inf, neginf=-np.inf)             # NaN ranks first
        ok, msg = o.topk_sets_match(ref, i2.cpu().numpy() - off, k, rtol=1e-5, atol=1e-4)
        assert ok, f"seg={seg}: {msg}"


def test_pairs_to_csr_kernel_matches_torch_path():
    """oov_pairs_to_csr (one CTA: count, scan, scatter, per-row sort) against the torch index plumbing, including
    padding rows (>= Q or negative), duplicate pairs, rows of 33-128 and of more than 128 entries, and an empty batch."""
    from oov_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(3)
    for Q, n, hot in ((1024, 25_600, 0), (1, 50, 0), (300, 5000, 200), (4096, 100_000, 150), (7, 3, 0), (64, 30_000, 0),
                      (8192, 51_200, 0), (20_000, 300_000, 0), (65_536, 1_500_000, 90)):   # beyond 8192 rows / 2^20 pairs: more CTAs, same kernel
        rows = torch.randint(-2, Q + 3, (n,), generator=g)
        if hot:
            rows[: hot] = 5 % Q                                      # one long row
        cols = torch.randint(0, 1_000_000, (n,), generator=g)
        cols[: n // 10] = cols[n // 10: 2 * (n // 10)]              # duplicates
        for cr in (None, ((1000, 300_000), (500_000, 777_777)), ((0, 0), (999_000, 1_000_000))):
            rp_k, c_k = ops.pairs_to_csr(rows.to(DEV), cols.to(DEV), Q, col_ranges=cr)
            rp_t, c_t = pu.pairs_to_csr_host(rows.to(DEV), cols.to(DEV), Q, col_ranges=cr)
            assert torch.equal(rp_k, rp_t), (Q, n, cr)
            m = int(rp_t[-1])
            assert torch.equal(c_k[:m], c_t[:m]), (Q, n, cr)
    rp, c = ops.pairs_to_csr(torch.zeros(0, dtype=torch.int64, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), 5)
    assert rp.tolist() == [0] * 6 and c.numel() == 0


def test_graphed_topk_equals_eager(tmp_path):
    """GraphedTopK (the whole step in one CUDA graph, padded history pairs) returns exactly what the eager
    model.full_sort_topk returns, for two different batches replayed through the same graph."""
    import cases
    import gpu_util as gu
    import oov_b200
    case = cases.CASES["bpr_lsh_ml100k"]
    inp = cases.retrieval_inputs(case)
    cfg, emb, model = gu.build_retrieval(case, inp, table_dtype="bfloat16")
    N = case.n_all_items
    Q, k = 64, 10
    gq = oov_b200.GraphedTopK(model, Q, k, N, max_pairs=Q * 8)
    g = torch.Generator(device="cpu").manual_seed(11)
    for trial in range(2):
        users = torch.randint(1, case.n_all_users, (Q,), generator=g).to(DEV)
        n_pairs = 100 + 200 * trial
        hu = torch.randint(0, Q, (n_pairs,), generator=g).to(DEV)
        hi = torch.randint(1, N, (n_pairs,), generator=g).to(DEV)
        s_g, i_g = gq(users, hu, hi)
        s_g, i_g = s_g.clone(), i_g.clone()
        s_e, i_e = model.full_sort_topk(users, k, n_total_items=N, history_index=(hu, hi))
        torch.cuda.synchronize()
        assert torch.equal(i_g, i_e) and torch.equal(s_g, s_e), trial


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_fused_table_equals_global(world):
    """ShardedRetrieval on one GPU, rank by rank (no process group): every rank scores its [in-vocab slice | OOV slice]
    table with ONE launch in local-row space (history rewritten by oov_pairs_to_csr's column map); merging the ranks'
    candidates reproduces the un-sharded full_sort_topk bit for bit, for the three collector segments."""
    import gpu_util as gu
    import oov_b200
    from oov_b200 import ops, sharded
    case = cases.CASES["bpr_lsh_ml100k"]
    inp = cases.retrieval_inputs(case)
    cfg, emb, model = gu.build_retrieval(case, inp, table_dtype="bfloat16")
    N, n_old = case.n_all_items, case.n_old_items
    Q, k = 96, 10
    g = torch.Generator(device="cpu").manual_seed(world)
    users = torch.randint(1, case.n_all_users, (Q,), generator=g).to(DEV)
    hu = torch.randint(0, Q, (900,), generator=g).to(DEV)
    hi = torch.randint(0, N, (900,), generator=g).to(DEV)
    user_e = model._assemble("user", users, out_dtype=model.table_dtype)
    for seg in ((0, 1 << 62), (0, n_old), (n_old, 1 << 62)):
        s_ref, i_ref = model.full_sort_topk(users, k, n_total_items=N, history_index=(hu, hi), seg=seg)
        cands, keys = [], []
        for r in range(world):
            sr = sharded.ShardedRetrieval(model, N, rank=r, world_size=world)
            assert sr.fused
            sr.build_shard()
            c = sr.fused_candidates(user_e, k, hist_pairs=(hu, hi), seg=seg)
            assert c is not None and c.shape == (1, Q, k, 2)
            cands.append(c)
            kk = sr.fused_keys(user_e, k, hist_pairs=(hu, hi), seg=seg)      # what the ranks all-gather: 8 bytes per candidate
            assert kk.shape == (Q, k) and kk.dtype == torch.int64
            keys.append(kk)
        cs, ci = sharded.unpack_candidates(torch.cat(cands, dim=0))
        s_m, i_m = ops.topk_merge(cs, ci)
        s_k, i_k = ops.topk_merge_keys(torch.stack(keys))
        torch.cuda.synchronize()
        assert torch.equal(i_m, i_ref) and torch.equal(s_m, s_ref), (world, seg)
        assert torch.equal(i_k, i_ref) and torch.equal(s_k, s_ref), (world, seg)


def test_tc_score_nan_rows_rank_first():
    """An all-zero LSH multi-hot row gives a NaN item embedding (lsh_embedder.py:158); torch.topk ranks NaN first."""
    from oov_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(1)
    users = torch.randn(40, 64, generator=g).to(torch.bfloat16).to(DEV)
    items = torch.randn(3000, 64, generator=g).to(torch.bfloat16).to(DEV)
    items[[5, 700, 2999]] = float("nan")
    s, i = ops.fullsort_topk(users, items, 5, mask_pad=False, path=ops.PATH_TCGEN05)
    assert (i[:, :3].cpu() == torch.tensor([5, 700, 2999])).all()
    assert torch.isnan(s[:, :3]).all() and not torch.isnan(s[:, 3:]).any()


def test_tc_score_shard_merge_equals_global():
    """Row-sharding property on the tensor-core path at a larger size (the multi-GPU data path, emulated as
    independent launches on one GPU): merge of per-shard top-ks == top-k over the whole table, bit for bit."""
    from oov_b200 import ops
    torch.manual_seed(0)
    Q, N, D, k = 512, 1_000_003, 64, 20
    users = torch.randn(Q, D, device=DEV).to(torch.bfloat16)
    items = torch.randn(N, D, device=DEV).to(torch.bfloat16)
    items[1000:1010] = items[900_000:900_010]
    hu = torch.randint(0, Q, (20_000,), device=DEV)
    hi = torch.randint(1, N, (20_000,), device=DEV)
    hist = ops.pairs_to_csr(hu, hi, Q)
    s_all, i_all = ops.fullsort_topk(users, items, k, hist=hist)
    bounds = [0, 250_000, 250_001, 777_777, N]
    cs, ci = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        s, i = ops.fullsort_topk(users, items[a:b], k, item_id_offset=a, hist=hist)
        cs.append(s)
        ci.append(i)
    ms, mi = ops.topk_merge(torch.stack(cs), torch.stack(ci))
    assert torch.equal(mi, i_all) and torch.equal(ms, s_all)
    ref = users.float() @ items.float().T
    ref[:, 0] = -float("inf")
    ref[hu, hi] = -float("inf")
    ok, msg = o.topk_sets_match(ref.cpu().numpy(), i_all.cpu().numpy(), k, rtol=1e-5, atol=1e-4)
    assert ok, msg


@pytest.mark.parametrize("n,F,B,D,dtype", [(1000, 32, 1000, 64, torch.float32), (70_001, 24, 1000, 64, torch.bfloat16),
                                           (513, 4, 100, 16, torch.float32), (300, 13, 3, 10, torch.float32),
                                           (20_000, 32, 129, 64, torch.float32),
                                           # F > 32: two 32-feature K chunks on the tensor cores (B' ring mode at B = 1000)
                                           (20_000, 64, 1000, 64, torch.float32), (30_001, 64, 1000, 64, torch.bfloat16),
                                           (777, 48, 300, 64, torch.float32), (2_000, 33, 1000, 32, torch.float32),
                                           (1_500, 64, 1100, 64, torch.float32), (900, 40, 500, 16, torch.bfloat16)])
def test_tc_lsh_matches_simt_bits_and_embeddings(n, F, B, D, dtype):
    """Tensor-core LSH (split-bf16 sign-projection GEMM + multi-hot x bucket-table GEMM) against the fp32 CUDA-core
    path: identical multi-hot bits (both resolve near-zero projections with the same fp32 FMA chain), embeddings within
    fp32 / bf16 tolerance, in-vocab rows gathered, all-zero rows NaN (lsh_embedder.py:158)."""
    from oov_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(n + F + B + D)
    n_old = n // 3
    feat = torch.nn.functional.normalize(torch.randn(n, F, generator=g), dim=-1).to(DEV)
    planes = torch.randn(B, F, generator=g).to(DEV)
    W = (torch.randn(B, D, generator=g) * 0.1).to(DEV)
    table = (torch.randn(n_old, D, generator=g) * 0.1).to(DEV)
    ids = torch.randperm(n, generator=g).to(DEV)              # in-vocab and OOV rows interleaved
    tc_t = torch.zeros(1, dtype=torch.int64, device=DEV)
    tc_s = torch.zeros(1, dtype=torch.int64, device=DEV)
    o_tc, b_tc = ops.lsh_embed(feat, planes, W, ids, out_dtype=dtype, n_old=n_old, iv_table=table, return_bits=True,
                               tie_count=tc_t, path=ops.PATH_TCGEN05)
    o_si, b_si = ops.lsh_embed(feat, planes, W, ids, out_dtype=dtype, n_old=n_old, iv_table=table, return_bits=True,
                               tie_count=tc_s, path=ops.PATH_SIMT_FP32)
    torch.cuda.synchronize()
    oov = (ids >= n_old)
    assert torch.equal(b_tc[oov], b_si[oov]), f"{(b_tc[oov] != b_si[oov]).sum().item()} words differ"
    assert int(tc_t.item()) == int(tc_s.item())
    a, b = o_tc.float(), o_si.float()
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=2 ** -7, atol=1e-4)
    assert torch.allclose(a[~torch.isnan(a)], b[~torch.isnan(b)], **tol), (a - b).abs().nan_to_num().max().item()
    assert torch.equal(o_tc[~oov], o_si[~oov])                 # in-vocab rows are plain copies
    # and against the oracle on a sample
    sel = torch.nonzero(oov).flatten()[:64].cpu().numpy()
    want = o.lsh_embed(feat.cpu().numpy(), ids.cpu().numpy()[sel], planes.cpu().numpy(), W.cpu().numpy())
    got = a.cpu().numpy()[sel]
    both = ~np.isnan(want)
    pu.assert_close(got[both], want[both], rtol=1e-4 if dtype == torch.float32 else 2 ** -6, atol=1e-5 if dtype == torch.float32 else 1e-3,
                    what="tc lsh vs oracle")


@pytest.mark.parametrize("n,F,B,D,dtype", [(40_000, 32, 1000, 64, torch.float32), (40_000, 32, 1000, 64, torch.bfloat16),
                                           (9_000, 32, 1500, 64, torch.bfloat16), (5_000, 20, 2100, 48, torch.float32),
                                           (3_000, 7, 37, 16, torch.float32),
                                           (40_000, 64, 1000, 64, torch.bfloat16), (6_000, 50, 2100, 64, torch.float32)])
def test_tc_lsh_deferred_sign_fix_matches_simt(n, F, B, D, dtype):
    """The product path of the tensor-core LSH (no multi-hot words requested): near-zero projections are queued, settled
    with the fp32 FMA chain at the end of the row tile and applied as rank-one corrections.  One wrong sign moves an
    fp32 output by ~2 |W| / count (1e-3 relative and more), so rtol 1e-5 against the CUDA-core path pins every bit;
    the reported tie counts must agree too.  Covers resident operands (B <= 1024, bf16), the bucket-table ring (fp32
    output: two pieces), both rings (B > 1024), unscaled features and degenerate rows (all-zero, NaN, Inf, tiny)."""
    from oov_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(7 * n + F + B + D)
    n_old = n // 4
    feat = torch.randn(n, F, generator=g)
    feat[n_old:n_old + n // 8] *= 1e4                                   # unnormalised rows ('none' normalisation)
    feat[n_old + n // 8:n_old + n // 4] *= 1e-9
    feat[n_old + 5] = 0.0                                              # every projection is an exact tie
    feat[n_old + 6, 1] = float("nan")
    feat[n_old + 7, 0] = float("inf")
    feat[n_old + 8] = 1e-30
    feat = feat.to(DEV)
    planes = torch.randn(B, F, generator=g)
    planes[min(3, B - 1)] *= 1e-3                                       # a short plane widens the tie band
    planes = planes.to(DEV)
    W = (torch.randn(B, D, generator=g) * 0.1).to(DEV)
    table = (torch.randn(n_old, D, generator=g) * 0.1).to(DEV)
    ids = torch.randperm(n, generator=g).to(DEV)
    tc_t = torch.zeros(1, dtype=torch.int64, device=DEV)
    tc_s = torch.zeros(1, dtype=torch.int64, device=DEV)
    o_tc = ops.lsh_embed(feat, planes, W, ids, out_dtype=dtype, n_old=n_old, iv_table=table, tie_count=tc_t, path=ops.PATH_TCGEN05)
    o_si = ops.lsh_embed(feat, planes, W, ids, out_dtype=dtype, n_old=n_old, iv_table=table, tie_count=tc_s, path=ops.PATH_SIMT_FP32)
    torch.cuda.synchronize()
    assert int(tc_t.item()) == int(tc_s.item()) and int(tc_t.item()) >= B      # the all-zero row alone gives B ties
    a, b = o_tc.float(), o_si.float()
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    ok = ~torch.isnan(a)
    tol = dict(rtol=1e-5, atol=2e-7) if dtype == torch.float32 else dict(rtol=2 ** -7, atol=1e-5)
    assert torch.allclose(a[ok], b[ok], **tol), (a - b).abs().nan_to_num().max().item()
    oov = ids >= n_old
    assert torch.equal(o_tc[~oov], o_si[~oov])


def test_dhe_memoised_hash_planes_equal_hashing_every_call(tmp_path):
    """dh_embedder.py:139 memoises the hashes of an id; here the byte planes of a known id range are kept by the embedder
    and later calls run the MLP alone.  Same bits as hashing on every call, also after the weights change (the planes
    depend on ids and keys only), in-vocab rows still gathered; the planes equal the generic uint32 hashes."""
    import gpu_util as G
    from oov_b200 import ops
    case = cases.DHE_CASES["dhe_scaled"]
    emb, keys, ws, bs = _dhe(case, tmp_path, ops.PATH_AUTO)
    lo, hi, n_old = 1000, 1000 + 30_011, 7000
    ids = torch.arange(lo, hi, device=DEV)
    table = (torch.randn(n_old, case.D, device=DEV) * 0.1).to(torch.bfloat16)
    for trial in range(2):
        a = emb.assemble_rows("item", ids, None, n_old, table, out_dtype=torch.bfloat16)
        b = emb.assemble_rows("item", ids, None, n_old, table, out_dtype=torch.bfloat16, id_range=(lo, hi))
        c = emb.assemble_rows("item", ids, None, n_old, table, out_dtype=torch.bfloat16, id_range=(lo, hi))     # from the cache
        torch.cuda.synchronize()
        assert len(emb._planes_cache) == 1
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)) and torch.equal(b.view(torch.int16), c.view(torch.int16))
        assert torch.equal(a[: n_old - lo], table[lo:n_old])
        with torch.no_grad():                                           # new weights, same planes
            emb.item_hash_net[2].weight.mul_(1.25)
    planes = next(iter(emb._planes_cache.values())).float()
    H = case.n_hashes
    h = ops.dhe_hash(ids, emb._keys_dev).to(torch.int64)
    assert torch.equal((planes[:, :H] * 65536 + planes[:, H:2 * H] * 256 + planes[:, 2 * H:3 * H]).to(torch.int64), h)


def test_slsh_single_bucket_and_gather_rows_out_validation():
    """n_buckets == 1: ceil(log2(1)) = 0 planes, every OOV id takes bucket 0 (single_lsh_embedder.py:77-87).
    gather_rows refuses an `out` of the wrong shape / stride / device instead of writing out of bounds."""
    from oov_b200 import ops
    feat = torch.randn(50, 6, device=DEV)
    planes = torch.zeros((0, 6), device=DEV)
    W = torch.randn(1, 8, device=DEV)
    ids = torch.arange(10, 40, device=DEV)
    out, buckets = ops.slsh_embed(feat, planes, 1, W, ids, return_buckets=True)
    torch.cuda.synchronize()
    assert (buckets == 0).all() and torch.equal(out, W.expand(30, 8))
    table = torch.randn(20, 8, device=DEV)
    idx = torch.arange(0, 20, 2, device=DEV)
    good = torch.empty((10, 8), device=DEV)
    assert torch.equal(ops.gather_rows(table, idx, out=good), table[idx])
    for bad in (torch.empty((9, 8), device=DEV), torch.empty((10, 7), device=DEV), torch.empty((10, 16), device=DEV)[:, ::2],
                torch.empty((10, 8))):
        with pytest.raises((ValueError, RuntimeError)):
            ops.gather_rows(table, idx, out=bad)


def test_lsh_side_cast_builds_the_same_item_table():
    """`build_item_rows_fused`: the fp32 -> bf16 cast of the in-vocab rows done by the LSH kernel's idle TMA warp
    (oov_lsh_embed_cast) gives the table of the two separate launches bit for bit — whole table, a row shard, sizes that
    are not multiples of the copy group, and the stand-alone cast when the OOV part is too small to fill the machine."""
    import oov_b200
    import gpu_util as G
    from oov_b200 import ops
    gen = torch.Generator().manual_seed(9)
    D, F_, B = 64, 32, 1000
    n_old, n_all = 30011, 30011 + 40037

    cfg = G.Config(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device=DEV, embedding_size=D, add_oov_buckets=True,
                   inductive_embedder="lsh", user_oov_buckets=B, item_oov_buckets=B, table_dtype="bfloat16", oov_normalization_type="global")
    uf = oov_b200.Interaction({"user_id": torch.arange(64), "f0": torch.randn(64, F_, generator=gen)})
    itf = oov_b200.Interaction({"item_id": torch.arange(n_all), "f0": torch.randn(n_all, F_, generator=gen)})
    ds = G.Dataset(32, n_old, uf, itf)
    emb = oov_b200.get_inductive_embedder(cfg, ds, mode="test-side-cast")
    model = oov_b200.BPR(cfg, ds, inductive_embedder=emb).to(DEV).eval()
    w = model.item_embedding.weight.detach()

    def separate(a, b, c, d):
        out = torch.empty((b - a + d - c, D), dtype=torch.bfloat16, device=DEV)
        ops.gather_rows(w, torch.arange(a, b, device=DEV), out=out[: b - a])
        emb.assemble_rows("item", torch.arange(c, d, device=DEV), model, 0, None, out=out[b - a:], out_dtype=torch.bfloat16)
        return out

    l0 = ops.launch_count()
    table = model.build_item_table(n_all)
    fused_launches = ops.launch_count() - l0
    want = separate(0, n_old, n_old, n_all)
    assert torch.equal(table.view(torch.int16), want.view(torch.int16))
    l0 = ops.launch_count()
    separate(0, n_old, n_old, n_all)
    assert fused_launches == ops.launch_count() - l0 - 1            # one launch fewer: the gather-cast is gone
    # a row shard [in-vocab slice | OOV slice] through the same entry
    a, b, c, d = 1000, 20001, n_old + 777, n_old + 777 + 25000
    out = torch.empty((b - a + d - c, D), dtype=torch.bfloat16, device=DEV)
    assert model.build_item_rows_fused((a, b), out[: b - a], (c, d), out[b - a:])
    assert torch.equal(out.view(torch.int16), separate(a, b, c, d).view(torch.int16))
    # few OOV rows (fewer row tiles than SMs): the cast runs as its own launch, same result
    a, b, c, d = 0, n_old, n_old, n_old + 300
    out = torch.empty((b - a + d - c, D), dtype=torch.bfloat16, device=DEV)
    assert model.build_item_rows_fused((a, b), out[: b - a], (c, d), out[b - a:])
    assert torch.equal(out.view(torch.int16), separate(a, b, c, d).view(torch.int16))
    # fp32 tables do not take the fused path
    model.table_dtype = torch.float32
    assert not model.build_item_rows_fused((0, 8), torch.empty((8, D), device=DEV), (n_old, n_old + 8), torch.empty((8, D), device=DEV))
