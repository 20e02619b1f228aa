"""Token gather + OOV overwrite of the ranking models — mirrors reference
model/abstract_recommender.py:715-842 (InductiveContextRecommender.embed_token_fields),
model/layers.py:130-153 (FMEmbedding) and :1617-1750 (InductiveFMFirstOrderLinear).

The gather/OOV-overwrite step feeds the dense towers; of those, DCNV2's cross network + MLP
(§8f row 2; dcnv2.py:120-144, 214-250) runs on the tensor-core linear here (class `DCNV2` below), as do the WideDeep
MLP (`WideDeep`) and xDeepFM's compressed interaction network + MLP (`xDeepFM`).
Column 0 of `token_fields` is the user id, column 1 the item id (abstract_recommender.py:691-692).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from torch import nn

from .. import ops


class FMEmbedding(nn.Module):
    """layers.py:130-153: one table for all token fields, per-field row offsets."""

    def __init__(self, field_dims: Sequence[int], offsets, embed_dim: int):
        super().__init__()
        self.embedding = nn.Embedding(int(sum(field_dims)), embed_dim)
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self._offsets_dev: Optional[torch.Tensor] = None

    def offsets_tensor(self, device) -> torch.Tensor:
        if self._offsets_dev is None or self._offsets_dev.device != torch.device(device):
            self._offsets_dev = torch.as_tensor(self.offsets, dtype=torch.int64, device=device)
        return self._offsets_dev

    def forward(self, input_x: torch.Tensor) -> torch.Tensor:
        w = self.embedding.weight.detach()
        return ops.token_gather(input_x, self.offsets_tensor(w.device), w, n_users=1 << 62, n_items=1 << 62)


class _TokenOOVMixin:
    """Shared OOV-overwrite logic of abstract_recommender.py:794-842 and layers.py:1634-1693."""

    def _embed_tokens(self, token_fields: torch.Tensor, uid_idx: int, iid_idx: int, out_dtype=None) -> torch.Tensor:
        w = self.token_embedding_table.embedding.weight.detach()
        dev = w.device
        token_fields = token_fields.to(dev)
        offsets = self.token_embedding_table.offsets_tensor(dev)
        fields = token_fields.shape[1]
        D = w.shape[1]
        emb, mapper = self.inductive_embedder, self.inductive_mapper
        if mapper is None and emb is None:
            raise RuntimeError("Must provide either self.inductive_mapper or self.inductive_embedder")
        # 1) every in-vocab cell; OOV user/item cells are left for step 2 (only the overwrite is observable)
        out = ops.token_gather(token_fields, offsets, w, self.n_users, self.n_items, uid_idx=uid_idx, iid_idx=iid_idx, out_dtype=out_dtype)
        # 2) OOV cells, written in place through a strided view (ids_stride = fields, out_stride = fields * D)
        for side, col, n_old in (("user", uid_idx, self.n_users), ("item", iid_idx, self.n_items)):
            ids = token_fields[:, col]
            view = out[:, col, :]
            if mapper is not None:
                mapped = mapper.map_user_ids(ids.contiguous()) if side == "user" else mapper.map_item_ids(ids.contiguous())
                buckets = (self.user_oov_buckets if side == "user" else self.item_oov_buckets).weight.detach()
                ops.gather_rows(buckets, mapped, idx_offset=-n_old, out=view)   # in-vocab ids (< n_old) are skipped
            else:
                emb.assemble_rows(side, ids, self, n_old, None, out=view, out_dtype=out.dtype)
        return out


class InductiveContextRecommender(nn.Module, _TokenOOVMixin):
    """The OOV-aware embedding front-end of DCNV2 / WideDeep / xDeepFM.

    `field_dims[0]` / `[1]` are the user / item vocabularies (n_users, n_items)."""

    def __init__(self, config, field_dims: Sequence[int], inductive_mapper=None, inductive_embedder=None,
                 first_order_embedder=None, first_order_mapper=None):
        super().__init__()
        self.embedding_size = config["embedding_size"]
        self.inductive_mapper = inductive_mapper
        self.inductive_embedder = inductive_embedder
        self.oov_training = False
        self.n_users, self.n_items = int(field_dims[0]), int(field_dims[1])
        if inductive_mapper is None and inductive_embedder is None:
            raise NotImplementedError("Must provide either self.inductive_mapper or self.inductive_embedder")
        offsets = np.array((0, *np.cumsum(field_dims)[:-1]), dtype=np.int64)
        self.token_field_offsets = offsets
        self.token_embedding_table = FMEmbedding(field_dims, offsets, self.embedding_size)
        try:
            add = config["add_oov_buckets"]
        except (KeyError, IndexError):
            add = False
        if add:
            self.n_user_oov_buckets = config["user_oov_buckets"]
            self.user_oov_buckets = nn.Embedding(self.n_user_oov_buckets, self.embedding_size)
            self.n_item_oov_buckets = config["item_oov_buckets"]
            self.item_oov_buckets = nn.Embedding(self.n_item_oov_buckets, self.embedding_size)
        if first_order_embedder is not None or first_order_mapper is not None:
            # abstract_recommender.py:748-760: a second, independently-initialised embedder with embedding_size = 1
            self.first_order_linear = InductiveFMFirstOrderLinear(config, field_dims, self.n_users, self.n_items,
                                                                  inductive_mapper=first_order_mapper,
                                                                  inductive_embedder=first_order_embedder)

    def embed_token_fields(self, token_fields: Optional[torch.Tensor], out_dtype=None) -> Optional[torch.Tensor]:
        """[B, fields] int64 -> [B, fields, D] (abstract_recommender.py:794-842); `out_dtype` = the dtype the consumer
        reads (default: the table's)."""
        if token_fields is None:
            return None
        return self._embed_tokens(token_fields, 0, 1, out_dtype=out_dtype)


class InductiveFMFirstOrderLinear(nn.Module, _TokenOOVMixin):
    """layers.py:1617-1750 restricted to the token fields: D = 1 table + D = 1 OOV buckets."""

    def __init__(self, config, field_dims: Sequence[int], n_users: int, n_items: int, output_dim: int = 1,
                 inductive_mapper=None, inductive_embedder=None):
        super().__init__()
        self.n_users, self.n_items = n_users, n_items
        self.inductive_mapper = inductive_mapper
        self.inductive_embedder = inductive_embedder
        offsets = np.array((0, *np.cumsum(field_dims)[:-1]), dtype=np.int64)
        self.token_field_offsets = offsets
        self.token_embedding_table = FMEmbedding(field_dims, offsets, output_dim)
        self.bias = nn.Parameter(torch.zeros((output_dim,)), requires_grad=True)
        try:
            add = config["add_oov_buckets"]
        except (KeyError, IndexError):
            add = False
        if add:
            self.n_user_oov_buckets = config["user_oov_buckets"]
            self.user_oov_buckets = nn.Embedding(self.n_user_oov_buckets, output_dim)
            self.n_item_oov_buckets = config["item_oov_buckets"]
            self.item_oov_buckets = nn.Embedding(self.n_item_oov_buckets, output_dim)

    def embed_token_fields(self, token_fields, uid_idx=None, iid_idx=None):
        """[B, fields] -> [B, 1, output_dim]: per-field first-order weights summed over fields."""
        if token_fields is None:
            return None
        w = self.token_embedding_table.embedding.weight.detach()
        dev = w.device
        token_fields = token_fields.to(dev)
        if w.shape[1] != 1 or uid_idx is None or iid_idx is None:
            if uid_idx is None or iid_idx is None:
                e = self.token_embedding_table(token_fields)
            else:
                e = self._embed_tokens(token_fields, uid_idx, iid_idx)
            return torch.sum(e, dim=1, keepdim=True)
        # D = 1 fast path: OOV cell values into two [B] scratch vectors, then one fused sum over fields
        offsets = self.token_embedding_table.offsets_tensor(dev)
        Bn = token_fields.shape[0]
        vals = {}
        for side, col, n_old in (("user", uid_idx, self.n_users), ("item", iid_idx, self.n_items)):
            ids = token_fields[:, col]
            scratch = torch.zeros((Bn, 1), dtype=torch.float32, device=dev)
            if self.inductive_mapper is not None:
                mp = self.inductive_mapper
                mapped = mp.map_user_ids(ids.contiguous()) if side == "user" else mp.map_item_ids(ids.contiguous())
                buckets = (self.user_oov_buckets if side == "user" else self.item_oov_buckets).weight.detach()
                ops.gather_rows(buckets, mapped, idx_offset=-n_old, out=scratch)
            elif self.inductive_embedder is not None:
                self.inductive_embedder.assemble_rows(side, ids, self, n_old, None, out=scratch, out_dtype=torch.float32)
            else:
                raise RuntimeError("Must provide either self.inductive_mapper or self.inductive_embedder")
            vals[side] = scratch.view(-1)
        s = ops.first_order_sum(token_fields, offsets, w, self.n_users, self.n_items, vals["user"], vals["item"],
                                uid_idx=uid_idx, iid_idx=iid_idx)
        return s.view(Bn, 1, 1)

    def forward(self, token_fields: torch.Tensor) -> torch.Tensor:
        """Token part of layers.py:1695-1750: sum over fields + bias -> [B, output_dim]."""
        return self.embed_token_fields(token_fields, 0, 1).sum(dim=1) + self.bias


class MLPLayers(nn.Module):
    """layers.py:33-92 with the DCNV2 settings (bn=True, activation='relu'): the same module tree, hence the same
    state_dict keys (`mlp_layers.{1,5,...}.weight`, `mlp_layers.{2,6,...}.running_mean`, ...)."""

    def __init__(self, layers: Sequence[int], dropout: float = 0.0, bn: bool = True):
        super().__init__()
        mods = []
        for i, o in zip(layers[:-1], layers[1:]):
            mods.append(nn.Dropout(p=dropout))
            mods.append(nn.Linear(i, o))
            if bn:
                mods.append(nn.BatchNorm1d(num_features=o))
            mods.append(nn.ReLU())
        self.mlp_layers = nn.Sequential(*mods)

    def folded(self):
        """[(W', b')] with every eval-mode BatchNorm folded into its Linear: W' = W gamma / sigma, b' = (b - mean) gamma / sigma + beta."""
        out, lin = [], None
        for m in self.mlp_layers:
            if isinstance(m, nn.Linear):
                lin = m
                out.append([m.weight.detach().float(), m.bias.detach().float()])
            elif isinstance(m, nn.BatchNorm1d):
                s = m.weight.detach().float() / torch.sqrt(m.running_var.float() + m.eps)
                w, b = out[-1]
                out[-1] = [w * s[:, None], (b - m.running_mean.float()) * s + m.bias.detach().float()]
        return out


class DCNV2(InductiveContextRecommender):
    """DCN-V2 ranking model on the OOV path (reference model/context_aware_recommender/dcnv2.py:30-250, `mixed: False`):
    token gather + OOV overwrite (`embed_token_fields`) feeding the cross network and the MLP, evaluated in eval mode on
    the tensor cores: every matrix product is `oov_tc_linear` (tcgen05, bf16 operands, fp32 accumulate; eval-mode
    BatchNorm folded into the Linear, ReLU / bias in the epilogue), the cross layers' x_0 * t + x_l tail is one
    elementwise kernel (`oov_cross_update`), the 1-wide predict layer + sigmoid one more linear.  Same parameter names as the
    reference (`cross_layer_w.{l}`, `bias.{l}`, `mlp_layers.mlp_layers.*`, `predict_layer.*`).  Token fields only
    (Criteo-shaped data: BASELINE configs[2]); float / token-sequence fields and training are outside this path."""

    def __init__(self, config, field_dims: Sequence[int], inductive_mapper=None, inductive_embedder=None):
        super().__init__(config, field_dims, inductive_mapper=inductive_mapper, inductive_embedder=inductive_embedder)

        def cfg(key, default):
            try:
                v = config[key]
            except (KeyError, IndexError):
                v = None
            return default if v is None else v

        if cfg("mixed", False):
            raise NotImplementedError("DCNV2 mixed (MoE low-rank) cross network is not on the accelerated path")
        self.structure = cfg("structure", "stacked")
        self.cross_layer_num = int(cfg("cross_layer_num", 3))
        self.mlp_hidden_size = list(cfg("mlp_hidden_size", [768, 768]))
        self.dropout_prob = float(cfg("dropout_prob", 0.2))
        self.num_feature_field = len(field_dims)
        self.in_feature_num = self.num_feature_field * self.embedding_size
        self.cross_layer_w = nn.ParameterList(nn.Parameter(torch.randn(self.in_feature_num, self.in_feature_num))
                                              for _ in range(self.cross_layer_num))
        self.bias = nn.ParameterList(nn.Parameter(torch.zeros(self.in_feature_num, 1)) for _ in range(self.cross_layer_num))
        self.mlp_layers = MLPLayers([self.in_feature_num] + self.mlp_hidden_size, dropout=self.dropout_prob, bn=True)
        top = self.mlp_hidden_size[-1] + (self.in_feature_num if self.structure == "parallel" else 0)
        self.predict_layer = nn.Linear(top, 1)
        for m in self.modules():                      # dcnv2.py:110 self.apply(xavier_normal_initialization)
            if isinstance(m, (nn.Embedding, nn.Linear)):
                nn.init.xavier_normal_(m.weight.data)
                if isinstance(m, nn.Linear) and m.bias is not None:
                    nn.init.constant_(m.bias.data, 0)
        self._packed = None

    # ---- weights in the layout the tensor-core linear takes (bf16 [N, K] with K padded to 8), rebuilt on demand
    def pack_tower(self):
        bf = torch.bfloat16

        def pad_k(w):
            k = w.shape[1]
            return torch.nn.functional.pad(w, (0, (-k) % 8)).to(bf).contiguous()

        pad = (-self.in_feature_num) % 8             # cross layers are square: pad rows like columns so x_l keeps x_0's padded width
        cross = [(pad_k(torch.nn.functional.pad(w.detach().float(), (0, 0, 0, pad))),
                  torch.nn.functional.pad(b.detach().float().reshape(-1), (0, pad)).contiguous()) for w, b in zip(self.cross_layer_w, self.bias)]
        mlp = [(pad_k(w), b.contiguous()) for w, b in self.mlp_layers.folded()]
        # the predict layer is 1 wide: pad its output dimension to 8 rows so it runs through the same kernel
        pw = self.predict_layer.weight.detach().float()
        pw8 = torch.zeros((8, pw.shape[1]), dtype=torch.float32, device=pw.device)
        pw8[0] = pw[0]
        pb8 = torch.zeros(8, dtype=torch.float32, device=pw.device)
        pb8[0] = self.predict_layer.bias.detach().float()[0]
        self._packed = dict(cross=cross, mlp=mlp, pred=(pad_k(pw8), pb8))
        return self._packed

    def cross_network(self, x0: torch.Tensor) -> torch.Tensor:
        """dcnv2.py:120-144 on bf16 rows [B, in_feature_num]."""
        pk = self._packed or self.pack_tower()
        xl = x0
        for w, b in pk["cross"]:
            # two launches per layer.  (A one-launch variant with the tail in the linear's store pass was measured on
            # B200 at in = 416, 65536 rows: 99 us against 43 + 27 us — the x_0 / x_l loads sit exposed in the epilogue.)
            t = ops.tc_linear(xl, w, b, act="none", out_dtype=torch.bfloat16)
            xl = ops.cross_update(x0, t, xl)
        return xl

    def _mlp(self, h: torch.Tensor) -> torch.Tensor:
        for w, b in (self._packed or self.pack_tower())["mlp"]:
            h = ops.tc_linear(h, w, b, act="relu", out_dtype=torch.bfloat16)
        return h

    def tower(self, x0: torch.Tensor) -> torch.Tensor:
        """[B, in_feature_num] bf16 embeddings -> [B] fp32 click probabilities (dcnv2.py:214-250, eval mode)."""
        if x0.shape[1] % 8:
            x0 = torch.nn.functional.pad(x0, (0, (-x0.shape[1]) % 8))
        pk = self._packed or self.pack_tower()
        cross = self.cross_network(x0)
        top = self._mlp(cross) if self.structure == "stacked" else torch.cat([cross[:, : self.in_feature_num], self._mlp(x0)], dim=1)
        pw, pb = pk["pred"]
        if self.structure != "stacked" and pw.shape[1] != top.shape[1]:
            top = torch.nn.functional.pad(top, (0, pw.shape[1] - top.shape[1]))
        return ops.tc_linear(top.contiguous(), pw, pb, act="sigmoid", out_dtype=torch.float32)[:, 0]

    def forward(self, interaction) -> torch.Tensor:
        tokens = interaction if isinstance(interaction, torch.Tensor) else interaction["token_fields"]
        emb = self.embed_token_fields(tokens, out_dtype=torch.bfloat16)     # [B, fields, D], written as bf16 by the gather
        return self.tower(emb.reshape(emb.shape[0], -1))

    def predict(self, interaction) -> torch.Tensor:
        return self.forward(interaction)


class WideDeep(InductiveContextRecommender):
    """Wide & Deep on the OOV path (reference model/context_aware_recommender/widedeep.py:33-100): the wide part is the
    D = 1 first-order linear over the token fields with its own OOV embedder (`InductiveFMFirstOrderLinear`,
    `oov_first_order_sum`), the deep part an MLP (Linear + ReLU, no BatchNorm: MLPLayers defaults) on the flattened
    `[B, fields * D]` embeddings and a 1-wide `deep_predict_layer`; `predict = sigmoid(wide + deep)`.  The MLP runs on
    `oov_tc_linear` (bf16 operands, fp32 accumulate, ReLU epilogue); the reference's default widths (32, 16, 8) are small
    GEMMs — the path is gather / HBM bound (BASELINE configs[3]).  Same parameter names as the reference
    (`mlp_layers.mlp_layers.*`, `deep_predict_layer.*`, `first_order_linear.*`).  Token fields only; eval mode."""

    def __init__(self, config, field_dims: Sequence[int], inductive_mapper=None, inductive_embedder=None,
                 first_order_embedder=None, first_order_mapper=None):
        super().__init__(config, field_dims, inductive_mapper=inductive_mapper, inductive_embedder=inductive_embedder,
                         first_order_embedder=first_order_embedder, first_order_mapper=first_order_mapper)
        if not hasattr(self, "first_order_linear"):
            raise NotImplementedError("WideDeep needs a first-order embedder or mapper for its wide part")
        try:
            hidden = config["mlp_hidden_size"]
        except (KeyError, IndexError):
            hidden = None
        self.mlp_hidden_size = list(hidden) if hidden is not None else [32, 16, 8]
        try:
            dp = config["dropout_prob"]
        except (KeyError, IndexError):
            dp = None
        self.dropout_prob = 0.1 if dp is None else float(dp)
        self.num_feature_field = len(field_dims)
        self.mlp_layers = MLPLayers([self.embedding_size * self.num_feature_field] + self.mlp_hidden_size, self.dropout_prob, bn=False)
        self.deep_predict_layer = nn.Linear(self.mlp_hidden_size[-1], 1)
        for m in self.modules():                                     # widedeep.py:52-60 _init_weights
            if isinstance(m, (nn.Embedding, nn.Linear)):
                nn.init.xavier_normal_(m.weight.data)
                if isinstance(m, nn.Linear) and m.bias is not None:
                    nn.init.constant_(m.bias.data, 0)
        self._packed = None

    def pack_tower(self):
        bf = torch.bfloat16

        def pad(w, rows=0):                                          # K to a multiple of 8, N up to `rows`
            w = torch.nn.functional.pad(w.detach().float(), (0, (-w.shape[1]) % 8, 0, max(0, rows - w.shape[0])))
            return w.to(bf).contiguous()

        layers = []
        for w, b in self.mlp_layers.folded():
            n8 = (-w.shape[0]) % 8                                   # the next layer's K must be a multiple of 8: pad this N with zero rows
            layers.append((pad(w, w.shape[0] + n8), torch.nn.functional.pad(b, (0, n8)).contiguous()))
        pw = pad(self.deep_predict_layer.weight, 8)
        pb = torch.nn.functional.pad(self.deep_predict_layer.bias.detach().float(), (0, 7)).contiguous()
        self._packed = dict(mlp=layers, pred=(pw, pb))
        return self._packed

    def deep(self, x0: torch.Tensor) -> torch.Tensor:
        """[B, fields * D] bf16 -> [B] fp32 deep logits."""
        pk = self._packed or self.pack_tower()
        if x0.shape[1] % 8:
            x0 = torch.nn.functional.pad(x0, (0, (-x0.shape[1]) % 8))
        h = x0.contiguous()
        for w, b in pk["mlp"]:
            h = ops.tc_linear(h, w, b, act="relu", out_dtype=torch.bfloat16)
        pw, pb = pk["pred"]
        return ops.tc_linear(h, pw, pb, act="none", out_dtype=torch.float32)[:, 0]

    def forward(self, interaction) -> torch.Tensor:
        tokens = interaction if isinstance(interaction, torch.Tensor) else interaction["token_fields"]
        emb = self.embed_token_fields(tokens, out_dtype=torch.bfloat16)
        wide = self.first_order_linear(tokens).reshape(-1)           # [B] fp32 (layers.py:1634-1693)
        return wide + self.deep(emb.reshape(emb.shape[0], -1))       # logits, widedeep.py:70-81

    def predict(self, interaction) -> torch.Tensor:
        return torch.sigmoid(self.forward(interaction))


class xDeepFM(InductiveContextRecommender):
    """xDeepFM on the OOV path (reference model/context_aware_recommender/xdeepfm.py:33-225): first-order linear with its
    own OOV embedder + compressed interaction network (CIN) + MLP on the gathered `[B, fields, D]` embeddings;
    `predict = sigmoid(first_order + cin_linear(CIN) + mlp)`.

    CIN layer k (xdeepfm.py:157-186): z = einsum("bhd,bmd->bhmd", X^{k-1}, X^0) viewed `[B, H*M, D]`, a kernel-size-1
    Conv1d over the channel axis, ReLU, then (direct = False) the first half of the channels feeds the next layer and
    the second half is sum-pooled over D into the output.  Here rows are (b, d) pairs: `oov_cin_outer` writes z as the
    bf16 A operand `[B*D, H*M]`, the Conv1d is `oov_tc_linear` (tcgen05, ReLU epilogue) whose output `[B*D, H_k]` is
    the next layer's X^k in place, and `oov_cin_pool_dot` folds the pooling with that layer's slice of `cin_linear`
    in fp32.  The batch is walked in chunks so z (B*D x 1300 bf16 at the default sizes) stays bounded.
    Same parameter names as the reference (`conv1d_list.{k}.*`, `mlp_layers.mlp_layers.*`, `cin_linear.*`,
    `first_order_linear.*`).  Token fields only; eval mode."""

    CIN_CHUNK = 8192          # batch rows per pass through the unfused CIN (z: 8192 * D * H*M * 2 bytes)
    fused_cin = True          # one kernel per CIN layer when the shapes allow (False: outer product + linear + pooling launches)

    def _cin_fusable(self, M: int) -> bool:
        hs = [M] + self.field_nums[1:-1]
        n_hid = [(s if self.direct else s // 2) for s in self.cin_layer_size[:-1]] + [0]
        return all(ops.cin_layer_supported(h, M, (s + 7) // 8 * 8, nh, (nh + 7) // 8 * 8)
                   for h, s, nh in zip(hs, self.cin_layer_size, n_hid))

    def __init__(self, config, field_dims: Sequence[int], inductive_mapper=None, inductive_embedder=None,
                 first_order_embedder=None, first_order_mapper=None):
        super().__init__(config, field_dims, inductive_mapper=inductive_mapper, inductive_embedder=inductive_embedder,
                         first_order_embedder=first_order_embedder, first_order_mapper=first_order_mapper)
        if not hasattr(self, "first_order_linear"):
            raise NotImplementedError("xDeepFM needs a first-order embedder or mapper for its linear part")

        def cfg(key, default):
            try:
                v = config[key]
            except (KeyError, IndexError):
                v = None
            return default if v is None else v

        self.mlp_hidden_size = list(cfg("mlp_hidden_size", [128, 128, 128]))
        self.dropout_prob = float(cfg("dropout_prob", 0.2))
        self.direct = bool(cfg("direct", False))
        self.cin_layer_size = list(cfg("cin_layer_size", [100, 100, 100]))
        if not self.direct:                                          # xdeepfm.py:50-57: even sizes when the output is split
            self.cin_layer_size = [int(x // 2 * 2) for x in self.cin_layer_size]
        self.num_feature_field = len(field_dims)
        self.conv1d_list = nn.ModuleList()
        self.field_nums = [self.num_feature_field]
        for layer_size in self.cin_layer_size:
            self.conv1d_list.append(nn.Conv1d(self.field_nums[-1] * self.field_nums[0], layer_size, 1))
            self.field_nums.append(layer_size if self.direct else layer_size // 2)
        self.mlp_layers = MLPLayers([self.embedding_size * self.num_feature_field] + self.mlp_hidden_size + [1], self.dropout_prob, bn=False)
        self.final_len = sum(self.cin_layer_size) if self.direct else sum(self.cin_layer_size[:-1]) // 2 + self.cin_layer_size[-1]
        self.cin_linear = nn.Linear(self.final_len, 1)
        for m in self.modules():                                     # xdeepfm.py:89-95 _init_weights
            if isinstance(m, (nn.Embedding, nn.Conv1d, nn.Linear)):
                nn.init.xavier_normal_(m.weight.data)
                if isinstance(m, nn.Linear) and m.bias is not None:
                    nn.init.constant_(m.bias.data, 0)
        self._packed = None

    def pack_tower(self):
        bf = torch.bfloat16

        def pad(w, rows=0):                                          # K to a multiple of 8, N up to `rows`
            w = torch.nn.functional.pad(w.detach().float(), (0, (-w.shape[1]) % 8, 0, max(0, rows - w.shape[0])))
            return w.to(bf).contiguous()

        cin, cin_fused = [], []
        for conv, h_prev in zip(self.conv1d_list, self.field_nums):
            w = conv.weight.detach()[:, :, 0]                         # [O, H*M, 1] -> [O, H*M]
            o8 = (w.shape[0] + 7) // 8 * 8
            cin.append((pad(w, o8), torch.nn.functional.pad(conv.bias.detach().float(), (0, o8 - w.shape[0])).contiguous()))
            # the fused kernel's channel layout (column h*Mp + m, Mp = fields rounded up to a power of two)
            cin_fused.append(ops.cin_pack_weight(w, h_prev, self.field_nums[0], o8) if self.field_nums[0] <= 64 else None)
        mlp = []
        for w, b in self.mlp_layers.folded():
            o8 = (w.shape[0] + 7) // 8 * 8
            mlp.append((pad(w, o8), torch.nn.functional.pad(b, (0, o8 - w.shape[0])).contiguous()))
        self._packed = dict(cin=cin, cin_fused=cin_fused, mlp=mlp, lin_w=self.cin_linear.weight.detach().float().reshape(-1).contiguous(),
                            lin_b=float(self.cin_linear.bias.detach().float()[0]))
        return self._packed

    def compressed_interaction_network(self, emb: torch.Tensor) -> torch.Tensor:
        """[B, fields, D] bf16 -> [B] fp32 = cin_linear(CIN(emb)) (xdeepfm.py:134-190, 198; activation ReLU)."""
        pk = self._packed or self.pack_tower()
        B, M, D = emb.shape
        last = len(self.cin_layer_size) - 1
        if self.fused_cin and self._cin_fusable(M):
            # one tcgen05 kernel per layer: the outer-product operand never leaves the SM (oov_cin_layer)
            out = torch.full((B,), pk["lin_b"], dtype=torch.float32, device=emb.device)
            x0t = emb.transpose(1, 2).contiguous().view(B * D, M)       # rows (b, d), fields along the row (34 MB at 65536 x 26 x 10)
            hidden, off = x0t, 0
            for i, ((_, b), w, size) in enumerate(zip(pk["cin"], pk["cin_fused"], self.cin_layer_size)):
                if self.direct:
                    n_hid, lo, n = size, 0, size
                elif i != last:
                    n_hid, lo, n = size // 2, size // 2, size // 2
                else:
                    n_hid, lo, n = 0, 0, size
                if i == last:
                    n_hid = 0                                         # nothing reads the last layer's hidden part
                hid = ops.cin_layer(hidden, x0t, B, D, w, b, n_hid, lo, n, pk["lin_w"][off: off + n], out)
                hidden = hid[:, :n_hid] if hid is not None else None
                off += n
            return out
        out = torch.empty((B,), dtype=torch.float32, device=emb.device)
        for r0 in range(0, B, self.CIN_CHUNK):
            x0 = emb[r0: r0 + self.CIN_CHUNK]
            bc = x0.shape[0]
            acc = out[r0: r0 + bc]
            hidden, off = x0, 0
            for i, ((w, b), size) in enumerate(zip(pk["cin"], self.cin_layer_size)):
                z = ops.cin_outer(hidden, x0, D, first=(i == 0))
                y = ops.tc_linear(z, w, b, act="relu", out_dtype=torch.bfloat16)       # [bc*D, size (padded to 8)]
                if self.direct:
                    lo, n, hidden = 0, size, y[:, :size]
                elif i != last:                                       # torch.split(output, 2 * [size // 2], 1): (next_hidden, direct)
                    lo, n, hidden = size // 2, size // 2, y[:, : size // 2]
                else:
                    lo, n = 0, size
                ops.cin_pool_dot(y, lo, n, bc, D, pk["lin_w"][off: off + n], pk["lin_b"] if i == 0 else 0.0, acc, accumulate=i > 0)
                off += n
        return out

    def deep(self, x0: torch.Tensor) -> torch.Tensor:
        """[B, fields * D] bf16 -> [B] fp32: MLPLayers(... + [1]) — ReLU after every Linear, the last one included."""
        pk = self._packed or self.pack_tower()
        if x0.shape[1] % 8:
            x0 = torch.nn.functional.pad(x0, (0, (-x0.shape[1]) % 8))
        h = x0.contiguous()
        for l, (w, b) in enumerate(pk["mlp"]):
            final = l == len(pk["mlp"]) - 1
            h = ops.tc_linear(h, w, b, act="relu", out_dtype=torch.float32 if final else torch.bfloat16)
        return h[:, 0]

    def forward(self, interaction) -> torch.Tensor:
        tokens = interaction if isinstance(interaction, torch.Tensor) else interaction["token_fields"]
        emb = self.embed_token_fields(tokens, out_dtype=torch.bfloat16)                # [B, fields, D]
        cin = self.compressed_interaction_network(emb)
        dnn = self.deep(emb.reshape(emb.shape[0], -1))
        return self.first_order_linear(tokens).reshape(-1) + cin + dnn                # xdeepfm.py:205

    def predict(self, interaction) -> torch.Tensor:
        return torch.sigmoid(self.forward(interaction))
