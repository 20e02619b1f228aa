"""ORACLE (test infrastructure, NOT product code): import the UNMODIFIED reference.

The reference (`/root/reference/RecBole`) is pure Python on top of torch.  It
imports a handful of pip packages that are not installed in this image; none of
them contributes arithmetic to the hot path except `csiphash` (SipHash-2-4, C),
which is replaced by `oracle/siphash24.c` (from-scratch restatement, KAT-pinned).

This module only works where `/root/reference` (or `$OOV_REFERENCE`) exists, i.e.
in the authoring container.  It is used by

* `tests/golden/make_golden.py` to produce the committed golden fixtures, and
* `tests/test_oracle_vs_reference.py` (CPU, skipped when the reference tree is
  absent) to check the numpy restatement in `oracle/oracle.py` against the live
  reference classes.

Nothing under the product package imports this file.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_REF_CANDIDATES = [
    os.environ.get("OOV_REFERENCE", ""),
    "/root/reference/RecBole",
]


def reference_root() -> str | None:
    for cand in _REF_CANDIDATES:
        if cand and os.path.isdir(os.path.join(cand, "recbole", "inductive")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def _install_stubs() -> None:
    # --- pyLSHash: only a default-argument type in torch_hash.py:6,32 -----------------
    class StorageBase:  # noqa: D401 - stub
        pass

    class InMemoryStorage(StorageBase):
        def __init__(self, *_a, **_k):
            self._d = {}

        def clear(self):
            self._d.clear()

    storage = _stub("pyLSHash.storage", StorageBase=StorageBase, InMemoryStorage=InMemoryStorage)
    _stub("pyLSHash", storage=storage)

    # --- csiphash: the one arithmetic dependency (dh_embedder.py:12) ------------------
    from oracle import oracle as _o

    def siphash24(key: bytes, msg: bytes) -> bytes:
        return _o.siphash24_bytes(key, msg)

    _stub("csiphash", siphash24=siphash24)

    # --- cosmetic / tooling packages -----------------------------------------------------
    _stub("scann")

    class _Any:
        def __getattr__(self, _name):
            return ""

    _stub("colorama", init=lambda *a, **k: None, Fore=_Any(), Style=_Any(), Back=_Any())

    import logging

    class ColoredFormatter(logging.Formatter):
        def __init__(self, fmt=None, datefmt=None, log_colors=None, **_k):
            super().__init__(fmt.replace("%(log_color)s", "") if fmt else fmt, datefmt)

    _stub("colorlog", ColoredFormatter=ColoredFormatter)

    class Texttable:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, _name):
            return lambda *a, **k: ""

    _stub("texttable", Texttable=Texttable)
    tune = _stub("ray.tune", report=lambda **k: None)
    _stub("ray", tune=tune)
    _stub("pyparsing", Optional=object)
    for name in ("thop", "hyperopt", "plotly"):
        try:
            __import__(name)
        except Exception:
            _stub(name)

    # --- numpy >= 2 removed aliases the reference still uses --------------------------
    for alias, target in (("long", np.int64), ("bool", np.bool_), ("int", np.int64),
                          ("float", np.float64), ("float_", np.float64),
                          ("complex_", np.complex128), ("unicode_", np.str_)):
        if not hasattr(np, alias):
            try:
                setattr(np, alias, target)
            except Exception:
                pass


_loaded = False


def load():
    """Install the stubs, put the reference on sys.path and return a namespace of the
    reference classes used by the golden generator / cross-check tests."""
    global _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set OOV_REFERENCE or mount /root/reference)")
    if not _loaded:
        _install_stubs()
        if root not in sys.path:
            sys.path.insert(0, root)
        _loaded = True

    ns = types.SimpleNamespace()
    from recbole.data.interaction import Interaction
    from recbole.inductive.torch_hash import TorchLSHash
    from recbole.inductive.lsh_embedder import LSHInductiveEmbedder
    from recbole.inductive.single_lsh_embedder import SingleLSHInductiveEmbedder
    from recbole.inductive.dh_embedder import DeepHashEmbedder
    from recbole.inductive.mean_embedder import MeanEmbedder
    from recbole.inductive.zero_embedder import ZeroEmbedder
    from recbole.inductive.feature_cache import InductiveFeatureCache
    from recbole.inductive.random_mapper import RandomOOVInductiveMapper
    from recbole.inductive import get_inductive
    from recbole.model.general_recommender.bpr import BPR
    from recbole.model.general_recommender.directau import DirectAU

    ns.Interaction = Interaction
    ns.TorchLSHash = TorchLSHash
    ns.LSHInductiveEmbedder = LSHInductiveEmbedder
    ns.SingleLSHInductiveEmbedder = SingleLSHInductiveEmbedder
    ns.DeepHashEmbedder = DeepHashEmbedder
    ns.MeanEmbedder = MeanEmbedder
    ns.ZeroEmbedder = ZeroEmbedder
    ns.InductiveFeatureCache = InductiveFeatureCache
    ns.RandomOOVInductiveMapper = RandomOOVInductiveMapper
    ns.get_inductive = get_inductive
    ns.BPR = BPR
    ns.DirectAU = DirectAU
    return ns


class RefConfig(dict):
    """dict whose missing keys read as None (mirrors configurator.py:583-584)."""

    def __getitem__(self, key):
        return dict.get(self, key, None)


class RefDataset:
    """Two-method stand-in for the RecBole Dataset the model constructors query."""

    def __init__(self, n_users: int, n_items: int, user_feat=None, item_feat=None,
                 uid_field="user_id", iid_field="item_id"):
        self._num = {uid_field: n_users, iid_field: n_items}
        self.user_num, self.item_num = n_users, n_items
        self._uf, self._if = user_feat, item_feat

    def num(self, field):
        return self._num[field]

    def get_user_feature(self):
        return self._uf

    def get_item_feature(self):
        return self._if
