"""Live pin of the oracle against the UNMODIFIED reference on FRESH seeds (not the committed cases).

Runs only where the reference tree exists (the authoring container: /root/reference or $OOV_REFERENCE);
skipped on the GPU box, where the committed fixtures under tests/golden/ do the pinning.  CPU only."""
import dataclasses
import importlib.util
import os
import sys

import numpy as np
import pytest

import cases
import parity_util as pu
from oracle import oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OOV_REFERENCE") or "/root/reference/RecBole"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "recbole")),
                                reason="reference tree not present (GPU box): golden fixtures pin the oracle instead")


@pytest.fixture(scope="module")
def gen():
    """tests/golden/make_golden.py imported as a module (it loads the reference through oracle/refshim.py)."""
    spec = importlib.util.spec_from_file_location("make_golden_live", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["make_golden_live"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("base,seed", [("directau_lsh_global", 90210), ("bpr_slsh_odd", 4242), ("bpr_mean", 77)])
def test_retrieval_live(gen, base, seed):
    case = dataclasses.replace(cases.CASES[base], name=f"live_{base}", seed=seed)
    inp = cases.retrieval_inputs(case)
    ref = gen.run_retrieval(case)                       # the reference classes themselves
    out = pu.oracle_retrieval(case, inp)                # the numpy restatement
    if case.embedder == "lsh":
        for side, B in (("user", case.B_user), ("item", case.B_item)):
            want = pu.unpack_bits(ref[f"{side}_bits"], B)
            ties = pu.tie_positions(ref[f"{side}_near_rows"], ref[f"{side}_near_cols"], ref[f"{side}_near_vals"])
            pu.check_bits(out[f"{side}_bits"], want, ties)
    if case.embedder == "slsh":
        for side in ("user", "item"):
            assert (out[f"{side}_bucket_ids"] == ref[f"{side}_bucket_ids"]).all()
    pu.assert_close(out["oov_user_emb"], ref["oov_user_emb"], what="oov_user_emb")
    pu.assert_close(out["oov_item_emb"], ref["oov_item_emb"], what="oov_item_emb")
    pu.assert_close(out["all_item_e"], ref["all_item_e"], what="all_item_e")
    scale = float(np.nanmax(np.abs(ref["scores_raw"][np.isfinite(ref["scores_raw"])])))
    pu.assert_close(out["scores_raw"], ref["scores_raw"], rtol=1e-5, atol=1e-5 * scale, what="scores_raw")
    ok, msg = o.topk_sets_match(out["scores_masked"], ref["topk_idx"], case.k, rtol=1e-5, atol=1e-6 * scale)
    assert ok, msg


def test_dhe_live(gen):
    base = cases.DHE_CASES["dhe_scaled_d16_h32"] if hasattr(cases, "DHE_CASES") else None
    if base is None:
        pytest.skip("no DHE case table")
    case = dataclasses.replace(base, name="live_dhe", seed=31337)
    ref = gen.run_dhe(case)
    keys = o.keys_to_array(cases.dhe_keys(case.seed, case.n_hashes))
    ids = cases.dhe_ids(case)
    assert (o.dhe_hashes(ids, keys) == ref["hashes"]).all()          # SipHash-2-4 mod 2^24: bit-exact
    ws, bs = cases.dhe_weights(case)
    pu.assert_close(o.dhe_embed(ids, keys, ws, bs), ref["item_emb"], rtol=1e-5, atol=1e-6, what="dhe item_emb")
