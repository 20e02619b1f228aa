"""Minimal stand-in for recbole.data.interaction.Interaction — just what the embedders and models
read: `.columns`, `[name]`, `len()` (rows), `.to(device)`.  A real RecBole Interaction works as well."""
from __future__ import annotations

from typing import Dict

import torch


class Interaction:
    def __init__(self, interaction: Dict[str, torch.Tensor]):
        self.interaction = {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(v)) for k, v in interaction.items()}
        self.length = -1
        for v in self.interaction.values():
            self.length = max(self.length, v.shape[0])

    @property
    def columns(self):
        return list(self.interaction.keys())

    def __getitem__(self, index):
        if isinstance(index, str):
            return self.interaction[index]
        return Interaction({k: v[index] for k, v in self.interaction.items()})

    def __contains__(self, item):
        return item in self.interaction

    def __len__(self):
        return self.length

    def to(self, device):
        return Interaction({k: v.to(device) for k, v in self.interaction.items()})
