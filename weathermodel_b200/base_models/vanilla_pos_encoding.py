"""Fixed sinusoidal position table (reference src/base_models/vanilla_pos_encoding.py:23-58). The buffer
`pos_encoding` [max_len, hidden_dim] is part of the state_dict; the add itself is fused into the CUDA
embedding kernel, `forward` exists for callers that use the module on its own."""
import math

import torch
import torch.nn as nn


class VanillaPositionalEncoding(nn.Module):
    pos_encoding: torch.Tensor

    def __init__(self, hidden_dim, max_len, device):
        assert hidden_dim % 2 == 0, "hidden_dim should be divisible by 2 for separate encoding"
        super().__init__()
        self.hidden_dim = hidden_dim
        table = torch.zeros(max_len, hidden_dim)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        freq = torch.exp(torch.arange(0, hidden_dim, 2).float() * (-math.log(10000.0) / hidden_dim))
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pos_encoding", table)

    def forward(self, token_embedding: torch.Tensor) -> torch.Tensor:
        _, seq_len, hidden_dim = token_embedding.shape
        if hidden_dim != self.hidden_dim:
            raise ValueError(f"hidden_dim mismatch: got {hidden_dim} != expected {self.hidden_dim}")
        return token_embedding + self.pos_encoding[:seq_len, :].unsqueeze(0)
