"""Golden fixture of the xDeepFM head: the UNMODIFIED reference code (xDeepFM.compressed_interaction_network / forward /
predict with nn.Conv1d, layers.MLPLayers in eval mode and cin_linear) run on seeded inputs.  Authoring container only
(needs /root/reference):
    python tests/golden/make_golden_xdeepfm.py   ->   tests/golden/xdeepfm_head.npz

The reference's xDeepFM constructor needs a full RecBole dataset; `forward` only reads `concat_embed_input_fields`,
`first_order_linear`, `conv1d_list`, `cin_layer_size`, `field_nums`, `direct`, `mlp_layers`, `cin_linear`, so it is
called on a bare object carrying exactly those attributes, built the way xdeepfm.py:43-86 builds them (the embedding /
first-order inputs are given; their own parity is covered by the context fixtures).  No reference code is copied."""
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
from recbole.model.context_aware_recommender.xdeepfm import xDeepFM  # noqa: E402
from recbole.model.layers import MLPLayers  # noqa: E402


def run(seed, batch, fields, D, cin_sizes, hidden, direct):
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn(batch, fields, D, generator=g) * 0.5
    fm = torch.randn(batch, 1, generator=g) * 0.5
    self = types.SimpleNamespace()
    self.concat_embed_input_fields = lambda interaction: emb
    self.first_order_linear = lambda interaction: fm
    self.direct = direct
    self.cin_layer_size = list(cin_sizes) if direct else [int(x // 2 * 2) for x in cin_sizes]     # xdeepfm.py:50-57
    self.conv1d_list = torch.nn.ModuleList()
    self.field_nums = [fields]
    for size in self.cin_layer_size:                                                               # xdeepfm.py:60-68
        self.conv1d_list.append(torch.nn.Conv1d(self.field_nums[-1] * self.field_nums[0], size, 1))
        self.field_nums.append(size if direct else size // 2)
    self.mlp_layers = MLPLayers([D * fields] + hidden + [1], dropout=0.2)                          # xdeepfm.py:71-74
    final_len = sum(self.cin_layer_size) if direct else sum(self.cin_layer_size[:-1]) // 2 + self.cin_layer_size[-1]
    self.cin_linear = torch.nn.Linear(final_len, 1)
    self.sigmoid = torch.nn.Sigmoid()
    with torch.no_grad():
        for m in list(self.conv1d_list) + [mm for mm in self.mlp_layers.modules() if isinstance(mm, torch.nn.Linear)] + [self.cin_linear]:
            fan_in = m.weight.shape[1]
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (1.5 / np.sqrt(fan_in)))
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    self.mlp_layers.eval()
    self.compressed_interaction_network = lambda x, activation="ReLU": xDeepFM.compressed_interaction_network(self, x, activation)
    self.forward = lambda interaction: xDeepFM.forward(self, interaction)
    with torch.no_grad():
        cin = xDeepFM.compressed_interaction_network(self, emb)                                    # xdeepfm.py:134-190
        logits = xDeepFM.forward(self, None)                                                       # xdeepfm.py:192-207
        prob = xDeepFM.predict(self, None)
    d = {"emb": emb.numpy(), "fm": fm.numpy(), "cin": cin.numpy(), "logits": logits.numpy(), "prob": prob.numpy(),
         "direct": np.array(int(direct)), "lin_w": self.cin_linear.weight.detach().numpy(), "lin_b": self.cin_linear.bias.detach().numpy()}
    for l, c in enumerate(self.conv1d_list):
        d[f"conv_w{l}"] = c.weight.detach().numpy()
        d[f"conv_b{l}"] = c.bias.detach().numpy()
    for l, li in enumerate(m for m in self.mlp_layers.modules() if isinstance(m, torch.nn.Linear)):
        d[f"mlp_w{l}"] = li.weight.detach().numpy()
        d[f"mlp_b{l}"] = li.bias.detach().numpy()
    return d


if __name__ == "__main__":
    out = {}
    for name, args in (("default", (61, 200, 26, 10, [100, 100, 100], [128, 128, 128], False)),
                       ("direct", (62, 130, 6, 16, [24, 17, 8], [64, 32], True)),
                       ("odd", (63, 77, 5, 8, [13, 10], [32], False))):
        for k, v in run(*args).items():
            out[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(HERE, "xdeepfm_head.npz"), **out)
    print("wrote xdeepfm_head.npz", {k: (v.shape, float(np.abs(v).mean())) for k, v in out.items() if k.endswith(".prob") or k.endswith(".cin")})
