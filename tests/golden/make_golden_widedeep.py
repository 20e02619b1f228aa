"""Golden fixture of the Wide & Deep head: the UNMODIFIED reference code (WideDeep.forward / predict with layers.MLPLayers
in eval mode and the deep_predict_layer) run on seeded inputs.  Authoring container only (needs /root/reference):
    python tests/golden/make_golden_widedeep.py   ->   tests/golden/widedeep_head.npz

The reference's WideDeep constructor needs a full RecBole dataset; `forward` only reads `concat_embed_input_fields`,
`first_order_linear`, `mlp_layers`, `deep_predict_layer`, so it is called on a bare object carrying exactly those
attributes (the embedding / first-order inputs are given; their own parity is covered by the context fixtures).  No
reference code is copied or modified."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.load()
from recbole.model.context_aware_recommender.widedeep import WideDeep  # noqa: E402
from recbole.model.layers import MLPLayers  # noqa: E402


def run(seed, batch, fields, D, hidden):
    g = torch.Generator().manual_seed(seed)
    in_dim = fields * D
    emb = torch.randn(batch, fields, D, generator=g) * 0.3
    fm = torch.randn(batch, 1, generator=g) * 0.5
    self = types.SimpleNamespace()
    self.concat_embed_input_fields = lambda interaction: emb
    self.first_order_linear = lambda interaction: fm
    self.mlp_layers = MLPLayers([in_dim] + hidden, 0.1)                # widedeep.py:44-47 (bn=False, relu)
    self.deep_predict_layer = torch.nn.Linear(hidden[-1], 1)
    self.sigmoid = torch.nn.Sigmoid()
    with torch.no_grad():
        for m in list(self.mlp_layers.modules()) + [self.deep_predict_layer]:
            if isinstance(m, torch.nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (1.0 / np.sqrt(m.in_features)))
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    self.mlp_layers.eval()
    self.forward = lambda interaction: WideDeep.forward(self, interaction)
    with torch.no_grad():
        logits = WideDeep.forward(self, None)                         # widedeep.py:70-81
        prob = WideDeep.predict(self, None)                           # widedeep.py:90-91
    d = {"emb": emb.numpy(), "fm": fm.numpy(), "logits": logits.numpy(), "prob": prob.numpy(),
         "pred_w": self.deep_predict_layer.weight.detach().numpy(), "pred_b": self.deep_predict_layer.bias.detach().numpy()}
    for l, li in enumerate(m for m in self.mlp_layers.modules() if isinstance(m, torch.nn.Linear)):
        d[f"mlp_w{l}"] = li.weight.detach().numpy()
        d[f"mlp_b{l}"] = li.bias.detach().numpy()
    return d


if __name__ == "__main__":
    out = {}
    for name, args in (("default", (51, 300, 26, 10, [32, 16, 8])), ("wide", (52, 130, 6, 16, [256, 128]))):
        for k, v in run(*args).items():
            out[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(HERE, "widedeep_head.npz"), **out)
    print("wrote widedeep_head.npz", {k: v.shape for k, v in out.items() if k.endswith(".prob")})
