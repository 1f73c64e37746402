"""WeatherBERT on the B200 encoder engine.

Same constructor, attributes, state_dict keys and forward signature as the reference class
(src/pretraining/models/weatherbert.py:13-121), so pickled checkpoints and trainers are interchangeable.
The stock nn.Linear / nn.TransformerEncoder modules are kept as PARAMETER CONTAINERS only (identical
initialisation RNG consumption and key names); `forward` never calls them -- it launches the sm_100a
kernel schedule of libwm_b200.so through weathermodel_b200.engine.
"""
import copy
from typing import Optional

import torch
import torch.nn as nn

from ...base_models.base_model import BaseModel
from ...base_models.vanilla_pos_encoding import VanillaPositionalEncoding
from ...engine import EncoderRuntime, encoder_apply
from ...utils.constants import MAX_CONTEXT_LENGTH


class WeatherBERT(BaseModel):
    def __init__(self, weather_dim, output_dim, device, num_heads=20, num_layers=8, hidden_dim_factor=24,
                 max_len=MAX_CONTEXT_LENGTH):
        super().__init__("weatherbert")
        self.weather_dim = weather_dim
        self.input_dim = weather_dim + 1 + 2  # weather + (year-1970)/100 + lat/360 + lon/180
        self.output_dim = output_dim
        self.max_len = max_len
        hidden_dim = hidden_dim_factor * num_heads
        self.in_proj = nn.Linear(self.input_dim, hidden_dim)
        self.positional_encoding = VanillaPositionalEncoding(hidden_dim, max_len=max_len, device=device)
        layer = nn.TransformerEncoderLayer(batch_first=True, d_model=hidden_dim, nhead=num_heads,
                                           dim_feedforward=hidden_dim * 4, device=device)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=num_layers)
        self.out_proj = nn.Linear(hidden_dim, output_dim)

    # the runtime holds device buffers and a C handle: never pickled, rebuilt lazily
    @property
    def runtime(self) -> EncoderRuntime:
        rt = self.__dict__.get("_wm_runtime")
        if rt is None:
            first = self.transformer_encoder.layers[0]
            rt = EncoderRuntime(self, num_heads=first.self_attn.num_heads, dropout_p=float(first.dropout.p),
                                ln_eps=float(first.norm1.eps))
            self.__dict__["_wm_runtime"] = rt
        first = self.transformer_encoder.layers[0]
        rt.dropout_p = float(first.dropout.p)  # follows nn.Dropout.p edits (dropout neutralisation in tests)
        return rt

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_wm_runtime", None)
        return state

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k != "_wm_runtime":
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def load_pretrained(self, pretrained_model: "WeatherBERT", load_out_proj: bool = True):
        """Deep-copy the encoder of a pretrained model (reference weatherbert.py:58-82)."""
        if self.input_dim != pretrained_model.input_dim:
            raise ValueError(f"expected input dimension {self.input_dim} but received {pretrained_model.input_dim}")
        if self.max_len != pretrained_model.max_len:
            raise ValueError(f"expected max length {self.max_len} but received {pretrained_model.max_len}")
        device = next(self.parameters()).device
        self.in_proj = copy.deepcopy(pretrained_model.in_proj).to(device)
        self.positional_encoding = copy.deepcopy(pretrained_model.positional_encoding).to(device)
        self.transformer_encoder = copy.deepcopy(pretrained_model.transformer_encoder).to(device)
        if load_out_proj:
            self.logger.info("Loading out_proj from pretrained model")
            self.out_proj = copy.deepcopy(pretrained_model.out_proj).to(device)
        else:
            self.logger.info("Not loading out_proj from pretrained model")
        self.__dict__.pop("_wm_runtime", None)  # parameters were replaced: re-flatten on next use

    def forward_raw(self, weather, coords, year, interval, weather_feature_mask, src_key_padding_mask=None):
        """Padded raw head output fp32 [B, S, 32|64] (columns >= out features are padding). `interval` is
        accepted and ignored exactly like the reference (normalised there and never used)."""
        if src_key_padding_mask is not None:
            raise NotImplementedError("src_key_padding_mask is always None on this path (SURVEY.md 3.5)")
        if weather.shape[1] > self.max_len:
            raise ValueError(f"sequence length {weather.shape[1]} exceeds max_len {self.max_len}")
        return encoder_apply(self.runtime, weather, coords, year, weather_feature_mask, self.training)

    def forward(self, weather: torch.Tensor, coords: torch.Tensor, year: torch.Tensor, interval: torch.Tensor,
                weather_feature_mask: torch.Tensor, src_key_padding_mask: Optional[torch.Tensor] = None):
        """weather [B,S,F] f32, coords [B,2] degrees, year [B,S], interval [B,1], mask [B,S,F] bool -> [B,S,out]"""
        y = self.forward_raw(weather, coords, year, interval, weather_feature_mask, src_key_padding_mask)
        return y[..., : self.out_proj.out_features]
