"""`inductive_mapper=random` — mirrors reference inductive/random_mapper.py:37-130 and
abstract_mapper.py: OOV id -> n_old + hash(id - n_old) % n_buckets with hash in
{mod, fast, 3round, 64bit}; int64 wrap-around, arithmetic shifts and Python-sign `%` are
reproduced bit-exactly by the `oov_map_ids` kernel.
"""
from __future__ import annotations

from torch import nn

from .. import ops


class AbstractInductiveMapper(nn.Module):
    def __init__(self, user_features, item_features) -> None:
        super().__init__()
        self.user_features = user_features
        self.item_features = item_features
        self.n_new_users = len(user_features)
        self.n_new_items = len(item_features)
        self.training = False

    def set_train(self):
        self.training = True

    def set_eval(self):
        self.training = False

    def map_user_ids(self, user_ids):
        raise NotImplementedError()

    def map_item_ids(self, item_ids):
        raise NotImplementedError()


class RandomOOVInductiveMapper(AbstractInductiveMapper):
    def __init__(self, user_features, item_features, n_original_users, n_original_items, n_user_oov_buckets,
                 n_item_oov_buckets, embedding_size, device, prime_pad, hash_function) -> None:
        super().__init__(user_features, item_features)
        self.n_original_users = n_original_users
        self.n_original_items = n_original_items
        self.n_user_oov_buckets = n_user_oov_buckets
        self.n_item_oov_buckets = n_item_oov_buckets
        self.embedding_size = embedding_size
        self.prime_pad = prime_pad
        self.hash_function = hash_function
        if hash_function not in ("mod", "fast", "3round", "64bit"):
            raise ValueError(f"Unknown hash function {hash_function}")

    def set_train(self):
        super().set_train()
        self.n_new_users = self.n_original_users * 2
        self.n_new_items = self.n_original_items * 2

    def set_eval(self):
        super().set_eval()
        self.n_new_users = len(self.user_features)
        self.n_new_items = len(self.item_features)

    def map_user_ids(self, user_ids):
        return ops.map_ids(user_ids, self.n_original_users, self.n_user_oov_buckets, self.hash_function)

    def map_item_ids(self, item_ids):
        return ops.map_ids(item_ids, self.n_original_items, self.n_item_oov_buckets, self.hash_function)
