"""Golden fixture of the DCN-V2 dense tower: the UNMODIFIED reference code (DCNV2.cross_network, layers.MLPLayers in eval
mode, the predict Linear + Sigmoid of DCNV2.forward) run on seeded inputs.  Authoring container only (needs
/root/reference):   python tests/golden/make_golden_dcnv2.py   ->   tests/golden/dcnv2_tower.npz

The reference's DCNV2 constructor needs a full RecBole dataset; the tower's maths lives in two methods that only read
`cross_layer_num`, `cross_layer_w`, `bias`, `mlp_layers`, `predict_layer`, `structure`, so they are called on a bare
object carrying exactly those attributes (no reference code is copied or modified)."""
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
from recbole.model.context_aware_recommender.dcnv2 import DCNV2  # noqa: E402
from recbole.model.layers import MLPLayers  # noqa: E402


def run(seed, batch, fields, D, hidden, structure):
    g = torch.Generator().manual_seed(seed)
    in_dim = fields * D
    self = types.SimpleNamespace()
    self.cross_layer_num = 3
    self.cross_layer_w = [torch.randn(in_dim, in_dim, generator=g) * (1.0 / np.sqrt(in_dim)) for _ in range(3)]
    self.bias = [torch.randn(in_dim, 1, generator=g) * 0.1 for _ in range(3)]
    self.mlp_layers = MLPLayers([in_dim] + hidden, dropout=0.2, bn=True)
    with torch.no_grad():
        for m in self.mlp_layers.modules():
            if isinstance(m, torch.nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (1.0 / np.sqrt(m.in_features)))
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            if isinstance(m, torch.nn.BatchNorm1d):               # "trained" running statistics and affine parameters
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    self.mlp_layers.eval()
    top_dim = hidden[-1] if structure == "stacked" else in_dim + hidden[-1]
    self.predict_layer = torch.nn.Linear(top_dim, 1)
    with torch.no_grad():
        self.predict_layer.weight.copy_(torch.randn(1, top_dim, generator=g) * (1.0 / np.sqrt(top_dim)))
        self.predict_layer.bias.fill_(0.05)
    x0 = torch.randn(batch, in_dim, generator=g) * 0.3
    with torch.no_grad():
        cross = DCNV2.cross_network(self, x0)                     # reference dcnv2.py:120-144
        if structure == "stacked":                                # reference dcnv2.py:236-246
            deep = self.mlp_layers(cross)
            out = torch.sigmoid(self.predict_layer(deep)).squeeze(1)
        else:                                                     # reference dcnv2.py:222-234
            deep = self.mlp_layers(x0)
            out = torch.sigmoid(self.predict_layer(torch.cat([cross, deep], dim=-1))).squeeze(1)
    d = {"x0": x0.numpy(), "cross_out": cross.numpy(), "deep_out": deep.numpy(), "out": out.numpy(),
         "pred_w": self.predict_layer.weight.detach().numpy(), "pred_b": self.predict_layer.bias.detach().numpy()}
    for l in range(3):
        d[f"cross_w{l}"] = self.cross_layer_w[l].numpy()
        d[f"cross_b{l}"] = self.bias[l].numpy()
    lins = [m for m in self.mlp_layers.modules() if isinstance(m, torch.nn.Linear)]
    bns = [m for m in self.mlp_layers.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    for l, (li, bn) in enumerate(zip(lins, bns)):
        d[f"mlp_w{l}"] = li.weight.detach().numpy(); d[f"mlp_b{l}"] = li.bias.detach().numpy()
        d[f"bn_mean{l}"] = bn.running_mean.numpy(); d[f"bn_var{l}"] = bn.running_var.numpy()
        d[f"bn_gamma{l}"] = bn.weight.detach().numpy(); d[f"bn_beta{l}"] = bn.bias.detach().numpy()
        d[f"bn_eps{l}"] = np.float32(bn.eps)
    return d


if __name__ == "__main__":
    out = {}
    for name, args in (("stacked", (41, 192, 6, 16, [96, 64], "stacked")), ("parallel", (42, 100, 5, 8, [64, 32], "parallel"))):
        for k, v in run(*args).items():
            out[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(HERE, "dcnv2_tower.npz"), **out)
    print("wrote dcnv2_tower.npz", {k: v.shape for k, v in out.items() if k.endswith(".out")})
