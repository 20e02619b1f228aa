"""Random-hyperplane LSH — mirrors reference inductive/torch_hash.py:10-60.

`uniform_planes` is an nn.ParameterList of `randn(hash_size, input_dim)` so checkpoints keep the
reference's key (`...lsh.uniform_planes.0`).  `hash_points` runs the sign-projection + warp-ballot
bit-pack kernel; the packed words are the native format, the {0,1} fp32 matrix the reference returns
is produced only on request.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class TorchLSHash(nn.Module):
    def __init__(self, hash_size, input_dim, num_hashtables=1, storage_instance=None, device="cuda"):
        super().__init__()
        self.hash_size = hash_size
        self.input_dim = input_dim
        self.num_hashtables = num_hashtables
        self.storage_instance = storage_instance
        self.device = device
        self.uniform_planes = nn.ParameterList([
            nn.Parameter(torch.randn(self.hash_size, self.input_dim, device=self.device))
            for _ in range(self.num_hashtables)])

    def hash_points_packed(self, planes: torch.Tensor, input_points: torch.Tensor, tie_count=None) -> torch.Tensor:
        """int32 words [n, ceil(hash_size/32)]; bit (b & 31) of word (b >> 5) = !(x . plane_b < 0)."""
        ids = torch.arange(input_points.shape[0], device=input_points.device)
        return ops.lsh_bits(input_points, planes, ids, tie_count=tie_count)

    def hash_points(self, planes: torch.Tensor, input_points: torch.Tensor) -> torch.Tensor:
        """fp32 {0,1} matrix [n, hash_size] like torch_hash.py:55-60 (R<0 -> 0, else 1)."""
        words = self.hash_points_packed(planes, input_points)
        shifts = torch.arange(32, device=words.device, dtype=torch.int32)
        bits = (words.unsqueeze(-1) >> shifts) & 1
        return bits.reshape(words.shape[0], -1)[:, : planes.shape[0]].to(torch.float32)
