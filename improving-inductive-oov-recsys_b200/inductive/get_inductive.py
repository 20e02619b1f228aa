"""Factory — mirrors reference inductive/get_inductive.py:16-138: same function names, keyword
set, config keys and the module-global feature cache keyed by `mode`.

Config keys honoured (properties/overall.yaml:59-119): inductive_embedder in {lsh, slsh, dhe, fdhe, dnn,
mean, zero}, inductive_mapper in {random}, user_oov_buckets, item_oov_buckets, embedding_size,
device, oov_prime_pad, oov_normalization_type, dhe_num_hashes, dhe_layer_size, oov_hash_function.
The reference's knn embedder (ScaNN, a third-party ANN library) is outside this path and raises.
"""
from __future__ import annotations

from typing import Union

from .abstract_embedder import AbstractInductiveEmbedder
from .dh_embedder import DeepHashEmbedder
from .dnn_embedder import DNNEmbedder
from .feat_dh_embedder import FeatDeepHashEmbedder
from .feature_cache import InductiveFeatureCache
from .lsh_embedder import LSHInductiveEmbedder
from .mean_embedder import MeanEmbedder
from .random_mapper import AbstractInductiveMapper, RandomOOVInductiveMapper
from .single_lsh_embedder import SingleLSHInductiveEmbedder
from .zero_embedder import ZeroEmbedder

feat_cache = InductiveFeatureCache()

OOV_PRIME_PAD_DEFAULT = 112062759511      # properties/overall.yaml:71


def _cfg(config, key, default=None):
    try:
        v = config[key]
    except (KeyError, IndexError):
        v = None
    return default if v is None else v


def get_inductive_mapper(config, dataset, user_num=None, item_num=None, embedding_size=None,
                         first_order=False) -> Union[AbstractInductiveMapper, None]:
    if embedding_size is None:
        embedding_size = config["embedding_size"]
    if _cfg(config, "inductive_mapper") == "random":
        return RandomOOVInductiveMapper(user_features=dataset.get_user_feature(),
                                        item_features=dataset.get_item_feature(),
                                        n_original_users=user_num or dataset.user_num,
                                        n_original_items=item_num or dataset.item_num,
                                        n_user_oov_buckets=config["user_oov_buckets"],
                                        n_item_oov_buckets=config["item_oov_buckets"],
                                        embedding_size=embedding_size,
                                        device=config["device"],
                                        prime_pad=_cfg(config, "oov_prime_pad", OOV_PRIME_PAD_DEFAULT),
                                        hash_function=_cfg(config, "oov_hash_function", "3round"))
    return None


def get_inductive_embedder(config, dataset, mode="transductive", user_num=None, item_num=None, embedding_size=None,
                           first_order=False) -> Union[AbstractInductiveEmbedder, None]:
    global feat_cache
    if feat_cache.get_mode() != mode:
        feat_cache = InductiveFeatureCache(mode=mode)       # reset when the mode changes (get_inductive.py:46-50)
    if embedding_size is None:
        embedding_size = config["embedding_size"]
    kind = _cfg(config, "inductive_embedder")
    common = dict(user_features=dataset.get_user_feature(),
                  item_features=dataset.get_item_feature(),
                  n_original_users=user_num or dataset.user_num,
                  n_original_items=item_num or dataset.item_num)
    buckets = dict(n_user_oov_buckets=_cfg(config, "user_oov_buckets"), n_item_oov_buckets=_cfg(config, "item_oov_buckets"))
    prime_pad = _cfg(config, "oov_prime_pad", OOV_PRIME_PAD_DEFAULT)
    norm = _cfg(config, "oov_normalization_type", "per-feature")
    device = config["device"]
    if kind == "lsh":
        return LSHInductiveEmbedder(**common, **buckets, embedding_size=embedding_size, device=device,
                                    prime_pad=prime_pad, normalization_type=norm, feature_cache=feat_cache)
    if kind == "slsh":
        return SingleLSHInductiveEmbedder(**common, **buckets, embedding_size=embedding_size, device=device,
                                          prime_pad=prime_pad, normalization_type=norm)
    if kind == "dhe":
        return DeepHashEmbedder(**common, **buckets, embedding_size=embedding_size, device=device,
                                prime_pad=prime_pad, num_hashes=_cfg(config, "dhe_num_hashes", 128))
    if kind == "mean":
        return MeanEmbedder(**common, **buckets, embedding_size=embedding_size, device=device)
    if kind == "zero":
        return ZeroEmbedder(**common, embedding_size=embedding_size, device=device)
    if kind == "fdhe":          # get_inductive.py:99-110
        return FeatDeepHashEmbedder(**common, **buckets, embedding_size=embedding_size, device=device, prime_pad=prime_pad,
                                    num_hashes=_cfg(config, "dhe_num_hashes", 128), dhe_layer_size=_cfg(config, "dhe_layer_size", 512))
    if kind == "dnn":           # get_inductive.py:111-121
        return DNNEmbedder(**common, **buckets, embedding_size=embedding_size, device=device, prime_pad=prime_pad,
                           dhe_layer_size=_cfg(config, "dhe_layer_size", 512))
    if kind == "knn":
        raise NotImplementedError(
            "inductive_embedder='knn' (ScaNN) is outside the accelerated path (lsh, slsh, dhe, fdhe, dnn, mean, zero)")
    return None
