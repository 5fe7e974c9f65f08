"""Fourier / sine positional encodings of xyz — mirror of the reference's models/position_embedding.py
(shift_scale_points :12-40, PositionEmbeddingCoordsSine :43-172).  `gauss_B` is a buffer drawn at construction
(:69-71), so it travels in the state dict and parity runs copy it."""
import math

import torch
from torch import nn


def shift_scale_points(pred_xyz, src_range, dst_range=None):
    """Affine map of xyz [B, N, 3] from src_range = [min [B,3], max [B,3]] to dst_range (default the unit cube)."""
    lo, hi = src_range
    if dst_range is None:
        dst_range = [torch.zeros((lo.shape[0], 3), device=lo.device), torch.ones((lo.shape[0], 3), device=lo.device)]
    assert lo.shape[0] == pred_xyz.shape[0] and lo.shape == hi.shape and lo.shape[-1] == pred_xyz.shape[-1]
    dlo, dhi = dst_range
    src_diff = hi[:, None, :] - lo[:, None, :]
    dst_diff = dhi[:, None, :] - dlo[:, None, :]
    return ((pred_xyz - lo[:, None, :]) * dst_diff) / src_diff + dlo[:, None, :]


class PositionEmbeddingCoordsSine(nn.Module):
    def __init__(self, temperature=10000, normalize=False, scale=None, pos_type="fourier", d_pos=None, d_in=3, gauss_scale=1.0):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        assert pos_type in ["sine", "fourier"]
        self.d_pos, self.temperature, self.normalize, self.pos_type = d_pos, temperature, normalize, pos_type
        self.scale = 2 * math.pi if scale is None else scale
        if pos_type == "fourier":
            assert d_pos is not None and d_pos % 2 == 0
            B = torch.empty((d_in, d_pos // 2)).normal_()
            B *= gauss_scale
            self.register_buffer("gauss_B", B)

    def get_sine_embeddings(self, xyz, num_channels, input_range):
        num_channels = self.d_pos
        xyz = xyz.clone()
        if self.normalize:
            xyz = shift_scale_points(xyz, src_range=input_range)
        ndim = num_channels // xyz.shape[2]
        if ndim % 2 != 0:
            ndim -= 1
        rems = num_channels - ndim * xyz.shape[2]
        embeds, prev_dim, dim_t = [], 0, None
        for d in range(xyz.shape[2]):
            cdim = ndim
            if rems > 0:
                cdim += 2
                rems -= 2
            if cdim != prev_dim:
                dim_t = torch.arange(cdim, dtype=torch.float32, device=xyz.device)
                dim_t = self.temperature ** (2 * (dim_t // 2) / cdim)
            raw = xyz[:, :, d]
            if self.scale:
                raw *= self.scale
            pos = raw[:, :, None] / dim_t
            embeds.append(torch.stack((pos[:, :, 0::2].sin(), pos[:, :, 1::2].cos()), dim=3).flatten(2))
            prev_dim = cdim
        return torch.cat(embeds, dim=2).permute(0, 2, 1)

    def get_fourier_embeddings(self, xyz, num_channels=None, input_range=None):
        if num_channels is None:
            num_channels = self.gauss_B.shape[1] * 2
        bsize, npoints = xyz.shape[0], xyz.shape[1]
        d_in, d_out = self.gauss_B.shape[0], num_channels // 2
        assert num_channels > 0 and num_channels % 2 == 0 and d_out <= self.gauss_B.shape[1] and d_in == xyz.shape[-1]
        xyz = xyz.clone()
        if self.normalize:
            xyz = shift_scale_points(xyz, src_range=input_range)
        xyz *= 2 * math.pi
        proj = torch.mm(xyz.view(-1, d_in), self.gauss_B[:, :d_out]).view(bsize, npoints, d_out)
        return torch.cat([proj.sin(), proj.cos()], dim=2).permute(0, 2, 1)  # [B, d_pos, N]

    def forward(self, xyz, num_channels=None, input_range=None):
        assert isinstance(xyz, torch.Tensor) and xyz.ndim == 3
        with torch.no_grad():
            if self.pos_type == "sine":
                return self.get_sine_embeddings(xyz, num_channels, input_range)
            return self.get_fourier_embeddings(xyz, num_channels, input_range)

    def extra_repr(self):
        return f"type={self.pos_type}, scale={self.scale}, normalize={self.normalize}"
