"""CPU: host-side logic of the product — feature-matrix build, factory / constructor surface and
state_dict keys (vs the golden fixtures of the reference), CSR helpers, shard arithmetic, metrics."""
import json
import os

import numpy as np
import pytest
import torch

import cases
import parity_util as pu


class Config(dict):
    def __getitem__(self, key):
        return dict.get(self, key, None)


class Dataset:
    def __init__(self, n_users, n_items, uf, itf):
        self._num = {"user_id": n_users, "item_id": n_items}
        self.user_num, self.item_num = n_users, n_items
        self._uf, self._if = uf, itf

    def num(self, f):
        return self._num[f]

    def get_user_feature(self):
        return self._uf

    def get_item_feature(self):
        return self._if


def _features(id_field, cols):
    import oov_b200
    d = {id_field: torch.arange(cols[0].shape[0])}
    for i, c in enumerate(cols):
        d[f"f{i}"] = torch.from_numpy(c)
    return oov_b200.Interaction(d)


def _cfg(case, kind, **kw):
    c = Config(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cpu",
               embedding_size=case.D, add_oov_buckets=True, inductive_embedder=kind,
               oov_normalization_type=getattr(case, "normalization", "per-feature"), topk=[10, 20])
    c.update(kw)
    return c


@pytest.mark.parametrize("name", ["bpr_lsh_ml100k", "directau_lsh_global", "bpr_lsh_tinybuckets", "directau_slsh", "bpr_slsh_odd"])
def test_feature_mats_and_state_dict_keys(name):
    import oov_b200
    case = cases.CASES[name]
    inp = cases.retrieval_inputs(case)
    g = pu.load_golden(name)
    ds = Dataset(case.n_old_users, case.n_old_items, _features("user_id", inp["user_cols"]), _features("item_id", inp["item_cols"]))
    cfg = _cfg(case, case.embedder, user_oov_buckets=case.B_user, item_oov_buckets=case.B_item)
    emb = oov_b200.get_inductive_embedder(cfg, ds, mode=f"host-{name}", user_num=case.n_old_users, item_num=case.n_old_items)
    assert emb.n_new_users == case.n_all_users and emb.n_new_items == case.n_all_items
    pu.assert_close(emb.user_feature_mat.numpy(), g["user_feature_mat"], rtol=1e-6, atol=1e-7, what="user_feature_mat")
    pu.assert_close(emb.item_feature_mat.numpy(), g["item_feature_mat"], rtol=1e-6, atol=1e-7, what="item_feature_mat")
    planes = emb.item_lsh.uniform_planes[0]
    want_planes = case.B_item if case.embedder == "lsh" else int(np.ceil(np.log2(case.B_item)))
    assert tuple(planes.shape) == (want_planes, g["item_feature_mat"].shape[1])
    cls = oov_b200.BPR if case.model == "BPR" else oov_b200.DirectAU
    model = cls(cfg, ds, inductive_mapper=None, inductive_embedder=emb)
    assert sorted(model.state_dict().keys()) == sorted(g["state_dict_keys"].tolist())
    assert model.n_new_items == case.n_all_items
    # training-mode switches propagate like abstract_recommender.py:147-171
    model.set_oov_train()
    assert emb.training and model.oov_training
    model.set_oov_eval()
    assert not emb.training


def test_model_requires_mapper_or_embedder():
    import oov_b200
    case = cases.CASES["bpr_mean"]
    ds = Dataset(10, 10, None, None)
    with pytest.raises(NotImplementedError):
        oov_b200.BPR(_cfg(case, None), ds)


def test_feature_cache_shared_between_lsh_embedders():
    """get_inductive.py:46-50 + lsh_embedder.py:77-106: second lsh embedder of the same mode reuses the
    cached matrices (how the first-order embedder shares features)."""
    import oov_b200
    case = cases.CASES["directau_lsh_global"]
    inp = cases.retrieval_inputs(case)
    ds = Dataset(case.n_old_users, case.n_old_items, _features("user_id", inp["user_cols"]), _features("item_id", inp["item_cols"]))
    cfg = _cfg(case, "lsh", user_oov_buckets=8, item_oov_buckets=8)
    a = oov_b200.get_inductive_embedder(cfg, ds, mode="cache-test")
    b = oov_b200.get_inductive_embedder(cfg, ds, mode="cache-test", embedding_size=1)
    assert a.item_feature_mat is b.item_feature_mat
    c = oov_b200.get_inductive_embedder(cfg, ds, mode="other-mode")
    assert c.item_feature_mat is not a.item_feature_mat


def test_factory_rejects_out_of_scope_embedders_and_unknown_norm():
    import oov_b200
    case = cases.CASES["bpr_mean"]
    inp = cases.retrieval_inputs(case)
    ds = Dataset(case.n_old_users, case.n_old_items, _features("user_id", inp["user_cols"]), _features("item_id", inp["item_cols"]))
    with pytest.raises(NotImplementedError):          # ScaNN-backed, third-party ANN: the one embedder left outside
        oov_b200.get_inductive_embedder(_cfg(case, "knn", user_oov_buckets=4, item_oov_buckets=4), ds, mode="x")
    assert oov_b200.get_inductive_embedder(_cfg(case, None), ds, mode="x") is None
    with pytest.raises(ValueError, match="Invalid normalization type"):
        oov_b200.get_inductive_embedder(_cfg(case, "slsh", user_oov_buckets=4, item_oov_buckets=4,
                                             oov_normalization_type="bogus"), ds, mode="y")


def test_dhe_key_file_protocol(tmp_path, monkeypatch):
    """dh_embedder.py:95-120: keys are created once under ./hash_keys/<n>.hashes and re-read afterwards.
    (Constructing the embedder on CPU touches no kernel; key upload needs a device, so stub it.)"""
    import oov_b200
    from oov_b200.inductive import dh_embedder
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(dh_embedder.ops, "keys_tensor", lambda keys, device: torch.zeros(len(keys), 16, dtype=torch.uint8))
    fu = oov_b200.Interaction({"user_id": torch.arange(4), "f0": torch.ones(4, 2)})
    a = dh_embedder.DeepHashEmbedder(fu, fu, 2, 2, 4, 4, 8, "cpu", cases.OOV_PRIME_PAD, 16)
    path = tmp_path / "hash_keys" / "16.hashes"
    assert path.exists()
    stored = json.load(open(path))
    assert len(stored) == 16 and all(len(bytes.fromhex(s)) == 16 for s in stored)
    b = dh_embedder.DeepHashEmbedder(fu, fu, 2, 2, 4, 4, 8, "cpu", cases.OOV_PRIME_PAD, 16)
    assert a.hash_keys == b.hash_keys
    want = {f"{s}_hash_net.{i}.{p}" for s in ("user", "item") for i in (0, 2, 4, 6) for p in ("weight", "bias")}
    assert want == set(a.state_dict().keys())
    assert a.item_hash_net[0].in_features == 16 and a.item_hash_net[6].out_features == 8


def test_pairs_to_csr():
    from oov_b200 import ops
    hu = torch.tensor([2, 0, 2, 2, 0])
    hi = torch.tensor([9, 4, 1, 5, 3])
    rp, cols = pu.pairs_to_csr_host(hu, hi, 4)
    assert rp.tolist() == [0, 2, 2, 5, 5] and cols.tolist() == [3, 4, 1, 5, 9]
    rp0, c0 = pu.pairs_to_csr_host(torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), 3)
    with pytest.raises(RuntimeError):                    # the product has no CPU path
        ops.pairs_to_csr(hu, hi, 4)
    assert rp0.tolist() == [0, 0, 0, 0] and c0.numel() == 0
    w_rp, w_c = pu.history_csr(hu.numpy(), hi.numpy(), 4)
    assert rp.tolist() == w_rp.tolist() and cols.tolist() == w_c.tolist()


def test_pairs_to_csr_column_map_and_padding():
    """Host restatement of oov_pairs_to_csr's contract: padding rows dropped, two-range column map to local rows."""
    from oov_b200 import ops
    hu = torch.tensor([2, 0, 2, 4, 2, 0, -1, 1])          # 4 and -1 are padding for Q = 4
    hi = torch.tensor([9, 4, 1, 7, 25, 3, 2, 30])
    rp, cols = pu.pairs_to_csr_host(hu, hi, 4)
    assert rp.tolist() == [0, 2, 3, 6, 6] and cols[:6].tolist() == [3, 4, 30, 1, 9, 25]
    # shard = items [0, 5) and [20, 40): local rows 0..4 and 5..24; items 7, 9 belong to other ranks
    rp, cols = pu.pairs_to_csr_host(hu, hi, 4, col_ranges=((0, 5), (20, 40)))
    assert rp.tolist() == [0, 2, 3, 5, 5] and cols[:5].tolist() == [3, 4, 15, 1, 10]


def test_sharded_local_segments():
    """ShardedRetrieval._local_seg: a global id filter as one range of local rows of [in-vocab slice | OOV slice]."""
    from oov_b200 import sharded

    class M:
        n_items = 100
    sr = sharded.ShardedRetrieval(M(), 260, rank=1, world_size=4, local_topk_fn=None, merge_fn=lambda a, b: (a, b))
    assert sr.segments == [(25, 50), (140, 180)] and sr.fused
    assert sr._local_seg((0, 1 << 62)) == (0, 65)
    assert sr._local_seg((0, 100)) == (0, 25)              # old items only
    assert sr._local_seg((100, 1 << 62)) == (25, 65)       # new items only
    assert sr._local_seg((30, 150)) == (5, 35)             # tail of range 0 + head of range 1: contiguous locally
    assert sr._local_seg((30, 45)) == (5, 20)
    assert sr._local_seg((0, 45)) == (0, 20)
    assert sr._local_seg((60, 120)) == (0, 0)              # nothing of this rank
    assert sr._local_seg((26, 40)) == (1, 15)
    assert sr._local_seg((10, 49)) == (0, 24) and sr._local_seg((30, 49)) == (5, 24)
    assert sr._local_seg((26, 175)) is not None and sr._local_seg((26, 175)) == (1, 60)
    # not contiguous in local rows: stops short of the end of range 0 but continues into range 1
    assert sr._local_seg((30, 49 + 100)) == (5, 34)        # 49 + 100 = 149 -> [30, 50) and [140, 149): contiguous
    assert sr._local_seg((0, 1 << 62)) == (0, 65)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_local_history_partitions_the_pairs(world):
    """Host logic of the fused shard path: every rank rewrites the history pairs to LOCAL rows of its
    [in-vocab slice | OOV slice] table and drops the others; mapping the local columns back to global ids and taking the
    union over the ranks must give back exactly the original pairs (each once)."""
    from oov_b200 import ops, sharded
    g = torch.Generator().manual_seed(world)
    Q, n_old, n_total = 37, 1000, 2600
    hu = torch.randint(0, Q, (900,), generator=g)
    hi = torch.randint(0, n_total, (900,), generator=g)
    want = sorted(zip(hu.tolist(), hi.tolist()))
    got = []
    for r in range(world):
        (lo0, hi0), (lo1, hi1) = sharded.shard_segments(n_old, n_total, r, world)
        rp, cols = pu.pairs_to_csr_host(hu, hi, Q, col_ranges=((lo0, hi0), (lo1, hi1)))
        n0 = hi0 - lo0
        for q in range(Q):
            row = cols[rp[q]:rp[q + 1]].tolist()
            assert row == sorted(row)                          # ascending per row: the scoring kernel's cursor relies on it
            for c in row:
                assert 0 <= c < n0 + (hi1 - lo1)
                got.append((q, c + lo0 if c < n0 else c - n0 + lo1))
    assert sorted(got) == want


def test_shard_arithmetic_and_packing():
    from oov_b200 import sharded
    for n, world in ((10, 4), (7, 8), (1_000_003, 8), (0, 2)):
        parts = [sharded.split_range(5, 5 + n, r, world) for r in range(world)]
        assert parts[0][0] == 5 and parts[-1][1] == 5 + n
        assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    segs = [sharded.shard_segments(100, 260, r, 4) for r in range(4)]
    assert sorted(x for s in segs for x in range(s[0][0], s[0][1])) == list(range(100))
    assert sorted(x for s in segs for x in range(s[1][0], s[1][1])) == list(range(100, 260))
    assert sharded.shard_segments(100, 260, 1, 4, balanced=False) == [(65, 130)]
    s = torch.tensor([[1.5, -0.0, float("-inf"), float("nan")]])
    i = torch.tensor([[7, 1 << 40, -1, 3]])
    s2, i2 = sharded.unpack_candidates(sharded.pack_candidates(s, i))
    assert torch.equal(i2, i) and torch.equal(s2.view(torch.int32), s.view(torch.int32))


def test_topk_metrics():
    from oov_b200.evaluator import topk_metrics
    rec = np.array([[1, 0, 0, 1, 2], [0, 0, 0, 0, 1], [0, 1, 0, 0, 3]])
    m = topk_metrics(rec, [2, 4])
    assert m["hit@2"] == pytest.approx(2 / 3) and m["hit@4"] == pytest.approx(2 / 3)
    assert m["recall@4"] == pytest.approx((1.0 + 0 + 1 / 3) / 3)
    assert m["precision@2"] == pytest.approx((0.5 + 0 + 0.5) / 3)
    assert m["mrr@4"] == pytest.approx((1 + 0 + 0.5) / 3)
    dcg0 = 1 + 1 / np.log2(5)
    idcg0 = 1 + 1 / np.log2(3)
    assert m["ndcg@4"] == pytest.approx((dcg0 / idcg0 + 0 + (1 / np.log2(3)) / (1 + 1 / np.log2(3) + 1 / np.log2(4))) / 3)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: oracle port of the reference step on the host cores) prints ONE JSON
    line with the keys the driver reads; a second rank of a torchrun launch prints nothing and exits 0."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "dhe100k", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "full_sort_topk_queries_per_s" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["config"]["workload"] == "dhe100k"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out2 = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert out2.returncode == 0 and not [ln for ln in out2.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.parametrize("name,slices", [("lsh1m", 8), ("dhe100k", 3)])
def test_bench_cpu_arm_strided_slices_equal_the_unsliced_batch(name, slices):
    """The CPU arm of bench.py walks the item axis in strided slices (items s, s + S, ...: every slice has the table's
    mix of in-vocab and OOV rows, so any number of timed steps is representative).  S slices with the running top-k
    merge give the lists of the unsliced batch (ids up to fp32 near-ties)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    wl = dict(b.WORKLOADS[name])
    wl.update(n_items=20_000, n_old_items=10_000, n_users=2_000, n_old_users=1_000, Q=48)
    one = b.CpuReference(wl, 1)
    assert one.step() is True
    many = b.CpuReference(wl, slices)
    done = [many.step() for _ in range(slices)]
    assert done == [False] * (slices - 1) + [True]
    v1, v2 = one.best[0].numpy(), many.best[0].numpy()
    np.testing.assert_allclose(v1, v2, rtol=1e-5, atol=1e-6)
    i1, i2 = one.best[1].numpy(), many.best[1].numpy()
    # the fp32 GEMM blocks a slice differently from the whole table: ids may only differ where neighbouring scores tie
    diff = i1 != i2
    if diff.any():
        gap = np.minimum(np.abs(np.diff(v1, axis=1, prepend=np.inf)), np.abs(np.diff(v1, axis=1, append=-np.inf)))
        assert (gap[diff] <= 1e-5 * np.abs(v1[diff]) + 1e-6).all()
        assert diff.mean() < 0.05


@pytest.mark.parametrize("name", list(cases.FEATNET_CASES))
def test_featnet_embedders_host_side(name, tmp_path):
    """fdhe / dnn constructors on the CPU: factory dispatch, feature matrices (per-column L2 normalisation, hstack) and
    state_dict keys against the reference-generated fixture; layer shapes follow dhe_layer_size / F / num_hashes."""
    import oov_b200
    case = cases.FEATNET_CASES[name]
    inp = cases.featnet_inputs(case)
    g = np.load(os.path.join(pu.GOLDEN_DIR, "featnet.npz"), allow_pickle=False)
    cwd = os.getcwd()
    os.chdir(tmp_path)                                  # ./hash_keys is CWD-relative (feat_dh_embedder.py:52)
    try:
        cfg = _cfg(case, case.kind, user_oov_buckets=4, item_oov_buckets=4, dhe_num_hashes=case.n_hashes, dhe_layer_size=case.layer)
        ds = Dataset(4, 4, _features("user_id", inp["user_cols"]), _features("item_id", inp["item_cols"]))
        emb = oov_b200.get_inductive_embedder(cfg, ds, mode=f"host-{name}", user_num=4, item_num=4)
        if case.kind == "fdhe":                        # the key file protocol: written once, read back identically
            emb2 = oov_b200.get_inductive_embedder(cfg, ds, mode=f"host-{name}", user_num=4, item_num=4)
            assert emb2.hash_keys == emb.hash_keys and len(emb.hash_keys) == case.n_hashes
            assert os.path.exists(os.path.join("hash_keys", f"{case.n_hashes}.hashes"))
    finally:
        os.chdir(cwd)
    assert type(emb).__name__ == ("FeatDeepHashEmbedder" if case.kind == "fdhe" else "DNNEmbedder")
    assert set(g[f"{name}.state_dict_keys"].tolist()) == set(emb.state_dict().keys())
    for side in ("user", "item"):
        pu.assert_close(getattr(emb, f"{side}_feature_mat").numpy(), g[f"{name}.{side}_feature_mat"], rtol=2e-6, atol=1e-7, what="feature_mat")
    F_ = int(sum(case.widths))
    H = case.n_hashes if case.kind == "fdhe" else 0
    assert tuple(emb.item_hash_net[0].weight.shape) == (case.layer, H + F_)
    assert tuple(emb.item_hash_net[6].weight.shape) == (case.D, case.layer)
    assert not emb.training
    emb.set_train()
    assert emb.training
    emb.set_eval()


def test_xdeepfm_widedeep_module_tree_and_cin_weight_layout():
    """Constructor surface of the ranking models on the CPU: parameter names of the reference module tree
    (xdeepfm.py:43-86, widedeep.py:40-50), CIN channel bookkeeping (field_nums, final_len, even sizes when the output is
    split), and the weight layout of the fused CIN kernel (column h*Mp + m)."""
    import oov_b200
    from oov_b200 import ops
    from oov_b200.inductive.zero_embedder import ZeroEmbedder
    z = lambda d: ZeroEmbedder(np.zeros((10, 1), np.float32), np.zeros((10, 1), np.float32), 40, 40, d, "cpu")
    cfg = {"embedding_size": 10, "mlp_hidden_size": [128, 128, 128], "dropout_prob": 0.2, "device": "cpu", "direct": False,
           "cin_layer_size": [100, 101, 100]}
    m = oov_b200.xDeepFM(cfg, [40, 40] + [50] * 24, inductive_embedder=z(10), first_order_embedder=z(1))
    assert m.cin_layer_size == [100, 100, 100] and m.field_nums == [26, 50, 50, 50] and m.final_len == 50 + 50 + 100
    keys = set(m.state_dict().keys())
    want = {"cin_linear.weight", "cin_linear.bias", "token_embedding_table.embedding.weight",
            "first_order_linear.token_embedding_table.embedding.weight", "first_order_linear.bias"}
    want |= {f"conv1d_list.{i}.{p}" for i in range(3) for p in ("weight", "bias")}
    want |= {f"mlp_layers.mlp_layers.{i}.{p}" for i in (1, 4, 7, 10) for p in ("weight", "bias")}      # Dropout, Linear, ReLU triples
    assert want <= keys, want - keys
    assert tuple(m.conv1d_list[1].weight.shape) == (100, 50 * 26, 1) and tuple(m.cin_linear.weight.shape) == (1, 200)
    assert tuple(m.mlp_layers.mlp_layers[10].weight.shape) == (1, 128)
    with pytest.raises(NotImplementedError):
        oov_b200.xDeepFM(cfg, [40, 40, 50], inductive_embedder=z(10))                  # no first-order embedder / mapper
    md = oov_b200.xDeepFM(dict(cfg, direct=True, cin_layer_size=[24, 17, 8]), [40, 40, 50, 50, 50, 50], inductive_embedder=z(10), first_order_embedder=z(1))
    assert md.cin_layer_size == [24, 17, 8] and md.field_nums == [6, 24, 17, 8] and md.final_len == 49
    w = oov_b200.WideDeep({"embedding_size": 10, "device": "cpu"}, [40, 40, 50], inductive_embedder=z(10), first_order_embedder=z(1))
    assert {"deep_predict_layer.weight", "mlp_layers.mlp_layers.1.weight", "mlp_layers.mlp_layers.7.bias"} <= set(w.state_dict().keys())
    assert w.mlp_hidden_size == [32, 16, 8]
    # fused-kernel weight layout: fields padded to a power of two, rows padded with zeros
    assert [ops.cin_field_pitch(v) for v in (2, 8, 9, 26, 32, 33, 64)] == [8, 8, 16, 32, 32, 64, 64]
    H, M, O = 3, 6, 5
    cw = torch.arange(O * H * M, dtype=torch.float32).reshape(O, H * M) + 1
    pk = ops.cin_pack_weight(cw, H, M, rows=8)
    assert pk.shape == (8, 24) and pk.dtype == torch.bfloat16
    for h in range(H):
        assert torch.equal(pk[:O, h * 8: h * 8 + M].float(), cw[:, h * M: (h + 1) * M])
        assert float(pk[:, h * 8 + M: (h + 1) * 8].abs().sum()) == 0.0
    assert float(pk[O:].abs().sum()) == 0.0
    # shapes the fused kernel takes (host-side predicate of the C-ABI; no GPU needed)
    assert ops.cin_layer_supported(50, 26, 104, 50, 56) and not ops.cin_layer_supported(50, 27, 104, 50, 56)
    assert not ops.cin_layer_supported(65, 26, 104, 50, 56) and not ops.cin_layer_supported(50, 26, 136, 50, 56)
    assert not ops.cin_layer_supported(24, 6, 24, 17, 24)
