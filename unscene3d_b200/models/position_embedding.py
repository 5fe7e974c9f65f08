"""Fourier positional encoding of xyz behind the reference's module interface (`PositionEmbeddingCoordsSine`,
models/position_embedding.py:43-172): same constructor, same `gauss_B` buffer (drawn at construction, :69-71, so it travels
in the state dict), same `forward(xyz [B, N, 3], num_channels, input_range) -> [B, d_pos, N]`.

The arithmetic (:128-160: shift / scale to the unit cube, 2 pi, projection on gauss_B, sin | cos) is one libus3d kernel per
scene, `us3d_fourier_posenc`, which writes rows [N, d_pos] — the layout every caller permutes the result into
(models/mask3d.py:195-196, 238-240); the returned [B, d_pos, N] tensor is a view of that.  `encode_rows` hands out the rows
directly.  The `sine` variant (:83-126) is not on the path of any shipped configuration (conf/model/mask3d.yaml:
positional_encoding_type "fourier") and is not provided.
"""
import math

import torch
from torch import nn


class PositionEmbeddingCoordsSine(nn.Module):
    fourier_core = None  # CPU tests install the oracle's restatement here; None = the CUDA kernel

    def __init__(self, temperature=10000, normalize=False, scale=None, pos_type="fourier", d_pos=None, d_in=3, gauss_scale=1.0):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if pos_type != "fourier":
            raise NotImplementedError(f"positional encoding '{pos_type}': only the Fourier features every shipped configuration "
                                      "uses are provided (models/position_embedding.py:128-160)")
        if d_pos is None or d_pos % 2 != 0 or d_in != 3:
            raise ValueError("Fourier features need an even d_pos and xyz input")
        self.d_pos, self.temperature, self.normalize, self.pos_type = d_pos, temperature, normalize, pos_type
        self.scale = 2 * math.pi if scale is None else scale
        self.register_buffer("gauss_B", torch.empty((d_in, d_pos // 2)).normal_() * gauss_scale)

    def encode_rows(self, xyz, lo=None, hi=None, num_channels=None):
        """xyz [N, 3] (+ the range [lo, hi] to normalise with when the module normalises) -> [N, num_channels] rows."""
        d_out = (self.gauss_B.shape[1] * 2 if num_channels is None else num_channels) // 2
        if d_out <= 0 or d_out > self.gauss_B.shape[1]:
            raise ValueError(f"num_channels must be an even number in [2, {2 * self.gauss_B.shape[1]}]")
        if not self.normalize:
            lo = hi = None
        elif lo is None or hi is None:
            raise ValueError("a normalising encoding needs input_range")
        core = type(self).fourier_core
        if core is None:
            from unscene3d_b200.engine import functional as Fn

            core = Fn.fourier_posenc
        with torch.no_grad():
            return core(xyz, self.gauss_B, d_out, lo, hi)

    def forward(self, xyz, num_channels=None, input_range=None):
        assert isinstance(xyz, torch.Tensor) and xyz.ndim == 3 and xyz.shape[-1] == 3
        lo, hi = (None, None) if input_range is None else input_range
        rows = [self.encode_rows(xyz[b], None if lo is None else lo[b], None if hi is None else hi[b], num_channels)
                for b in range(xyz.shape[0])]
        return torch.stack(rows).permute(0, 2, 1)  # [B, d_pos, N]

    def extra_repr(self):
        return f"type={self.pos_type}, scale={self.scale}, normalize={self.normalize}"
