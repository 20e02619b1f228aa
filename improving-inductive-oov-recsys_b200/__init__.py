"""B200-native OOV inductive-embedding + full-sort top-k path (drop-in for the RecBole
`inductive_embedder` plugin API of snap-research/improving-inductive-oov-recsys).

Layout
    csrc/ + liboov_b200.so   hand-written sm_100a kernels behind the C-ABI in include/oov_b200.h
    _lib.py / ops.py         ctypes binding and tensor-level wrappers (no CPU fallback)
    inductive/               mirror of recbole/inductive: lsh, slsh, dhe, fdhe, dnn, mean, zero embedders,
                             random mapper, get_inductive_embedder / get_inductive_mapper
    model/                   BPR / DirectAU (fused assemble + full_sort_topk), context token gather
    evaluator.py             InductiveEvaluator / Collector on the fused path
    sharded.py               row-sharded retrieval, NCCL all-gather top-k merge
    graphed.py               the whole retrieval step captured in one CUDA graph (static shapes)

The directory name contains '-', so import it as `import oov_b200` (alias module at the repo root)
or `importlib.import_module("improving-inductive-oov-recsys_b200")`.
"""
from . import _lib, ops, sharded, graphed, evaluator, interaction        # noqa: F401
from .interaction import Interaction                             # noqa: F401
from .inductive import (abstract_embedder, feature_cache, torch_hash, lsh_embedder, single_lsh_embedder,   # noqa: F401
                        dh_embedder, feat_dh_embedder, dnn_embedder, mean_embedder, zero_embedder, random_mapper,
                        get_inductive)
from .model import general, context                              # noqa: F401
from .inductive.get_inductive import get_inductive_embedder, get_inductive_mapper   # noqa: F401
from .model.general import BPR, DirectAU                         # noqa: F401
from .model.context import DCNV2, WideDeep, xDeepFM, InductiveContextRecommender    # noqa: F401
from .evaluator import InductiveEvaluator, Collector            # noqa: F401
from .graphed import GraphedTopK, GraphedRanker                                 # noqa: F401

__version__ = "0.1.0"
