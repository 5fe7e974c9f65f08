"""Mask3D transformer mask-decoder — host-side mirror of the reference's models/mask3d.py (Mask3D :16-446,
SelfAttentionLayer :491-546, CrossAttentionLayer :548-609, FFNLayer :611-651) and models/modules/helpers_3detr.py
(GenericMLP :45-112).

Same constructor arguments, attribute / state-dict names and forward semantics:

    backbone (Res16UNet, out_fpn) -> 5 feature maps                                      models/mask3d.py:201
    raw-xyz pyramid by 4x MinkowskiAvgPooling on the UNet's own coordinate maps          :205-215
    Fourier pos-enc per level and scene                                                 :183-198, 217
    mask features = 1x1 conv, per-segment mean (scatter_mean)                           :218-223
    queries = FPS(voxel coords, num_queries) -> pos-enc -> GenericMLP                    :227-249
    num_decoders x len(hlevels) rounds of: mask_module -> sampled masked cross-attention -> self-attention -> FFN
    final mask_module                                                                   :271-395

The sparse operators, FPS and scatter_mean come from whatever is importable as `MinkowskiEngine`, `pointnet2._ext`
and `torch_scatter` (in the product: the sm_100a kernels behind unscene3d_b200/shims); attention layers are
torch.nn.MultiheadAttention exactly as in the reference.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.nn import functional as F

import MinkowskiEngine.MinkowskiOps as me
from MinkowskiEngine.MinkowskiPooling import MinkowskiAvgPooling
import pointnet2._ext as _pointnet2_ext
from torch_scatter import scatter_max, scatter_mean

from .modules.common import conv
from .position_embedding import PositionEmbeddingCoordsSine


def furthest_point_sample(xyz, npoint):
    """[B, N, 3] float -> int32 [B, npoint] (third_party/pointnet2/pointnet2_utils.py:22-30)."""
    return _pointnet2_ext.furthest_point_sampling(xyz.contiguous(), npoint)


class GenericMLP(nn.Module):
    """Conv1d/Linear stack; only the pieces Mask3D configures (helpers_3detr.py:45-112)."""

    def __init__(self, input_dim, hidden_dims, output_dim, norm_fn_name=None, activation="relu", use_conv=False,
                 dropout=None, hidden_use_bias=False, output_use_bias=True, output_use_activation=False,
                 output_use_norm=False, weight_init_name=None):
        super().__init__()
        assert norm_fn_name is None and activation == "relu" and dropout is None and not output_use_norm
        make = (lambda i, o, b: nn.Conv1d(i, o, 1, bias=b)) if use_conv else (lambda i, o, b: nn.Linear(i, o, bias=b))
        layers, prev = [], input_dim
        for width in hidden_dims:
            layers += [make(prev, width, hidden_use_bias), nn.ReLU()]
            prev = width
        layers.append(make(prev, output_dim, output_use_bias))
        if output_use_activation:
            layers.append(nn.ReLU())
        self.layers = nn.Sequential(*layers)
        if weight_init_name == "xavier_uniform":
            for p in self.parameters():
                if p.dim() > 1:
                    nn.init.xavier_uniform_(p)

    def forward(self, x):
        return self.layers(x)


def _activation(name):
    if name == "relu":
        return F.relu
    if name == "gelu":
        return F.gelu
    if name == "glu":
        return F.glu
    raise RuntimeError(f"activation should be relu/gelu, not {name}.")


class _DecoderLayer(nn.Module):
    def __init__(self, d_model, dropout, activation, normalize_before):
        super().__init__()
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.activation = _activation(activation)
        self.normalize_before = normalize_before

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos


class SelfAttentionLayer(_DecoderLayer):
    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        nn.Module.__init__(self)
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.activation = _activation(activation)
        self.normalize_before = normalize_before
        self._reset_parameters()

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        src = self.norm(tgt) if self.normalize_before else tgt
        q = k = self.with_pos_embed(src, query_pos)
        out = self.self_attn(q, k, value=src, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)[0]
        tgt = tgt + self.dropout(out)
        return tgt if self.normalize_before else self.norm(tgt)


class HeadSharedMask:
    """[B, K, Q] bool attention mask (True = hidden), the same for every head — the decoder's tensor before the
    reference's `repeat_interleave(num_heads, dim=0).permute((0, 2, 1))` (models/mask3d.py:358)."""

    def __init__(self, bkq):
        self.bkq = bkq

    def torch_layout(self, num_heads):
        return self.bkq.repeat_interleave(num_heads, dim=0).permute((0, 2, 1))


def _cuda_segment_attention_core(model, mask_features, output_segments, point2segment, num_pooling_steps):
    from unscene3d_b200.engine import functional as Fn  # CUDA only: there is no CPU path

    key, bits = Fn.segment_attention_masks(mask_features, output_segments, point2segment, num_pooling_steps)
    return me.SparseTensor(features=bits, coordinate_manager=mask_features.coordinate_manager, coordinate_map_key=key)


def _cuda_prepare_segment_attention(x, point2segment, max_steps):
    from unscene3d_b200.engine import functional as Fn

    Fn.prepare_segment_attention(x, point2segment, max_steps)


def _cuda_attention_core(mha, query, key, value, attn_mask=None):
    from unscene3d_b200.engine import functional as Fn  # CUDA only: there is no CPU path

    return Fn.multihead_cross_attention(mha, query, key, value, attn_mask=attn_mask)


class CrossAttentionLayer(_DecoderLayer):
    # The masked attention core is the libus3d kernel (csrc/attention.cu).  Like the sparse operators, which these model
    # files take from whatever is importable as `MinkowskiEngine`, it can be swapped: the CPU tests that run these
    # definitions over the oracle install the oracle's restatement here (tests/helpers.py).
    attention_core = None

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        nn.Module.__init__(self)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.activation = _activation(activation)
        self.normalize_before = normalize_before
        self._reset_parameters()

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        src = self.norm(tgt) if self.normalize_before else tgt
        if memory_key_padding_mask is not None:
            raise RuntimeError("CrossAttentionLayer: memory_key_padding_mask is not used by the reference (models/mask3d.py:358)")
        core = type(self).attention_core or _cuda_attention_core
        out = core(self.multihead_attn, self.with_pos_embed(src, query_pos), self.with_pos_embed(memory, pos), memory,
                   attn_mask=memory_mask)
        tgt = tgt + self.dropout(out)
        return tgt if self.normalize_before else self.norm(tgt)


class FFNLayer(_DecoderLayer):
    def __init__(self, d_model, dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False):
        nn.Module.__init__(self)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        self.activation = _activation(activation)
        self.normalize_before = normalize_before
        self._reset_parameters()

    def forward(self, tgt):
        src = self.norm(tgt) if self.normalize_before else tgt
        out = self.linear2(self.dropout(self.activation(self.linear1(src))))
        tgt = tgt + self.dropout(out)
        return tgt if self.normalize_before else self.norm(tgt)


class Mask3D(nn.Module):
    # Attention mask of a decoder round from the SEGMENT logits (train_on_segments): the libus3d sparse product
    # (engine.functional.segment_attention_masks).  Swappable like the sparse operators: the CPU tests that run this file over
    # the oracle install the reference's sequence (gather to the voxels, concatenate, pool, threshold) from oracle/ops_cpu.py.
    segment_attention_core = None
    def __init__(self, config, hidden_dim, num_queries, num_heads, dim_feedforward, sample_sizes, shared_decoder,
                 num_classes, num_decoders, dropout, pre_norm, positional_encoding_type, non_parametric_queries,
                 train_on_segments, normalize_pos_enc, use_level_embed, scatter_type, hlevels, use_np_features,
                 voxel_size, max_sample_size, random_queries, gauss_scale, random_query_both, random_normal):
        super().__init__()
        self.random_normal, self.random_query_both, self.random_queries = random_normal, random_query_both, random_queries
        self.max_sample_size, self.gauss_scale, self.voxel_size = max_sample_size, gauss_scale, voxel_size
        self.scatter_type, self.hlevels, self.use_level_embed = scatter_type, hlevels, use_level_embed
        self.train_on_segments, self.normalize_pos_enc = train_on_segments, normalize_pos_enc
        self.num_decoders, self.num_classes, self.dropout, self.pre_norm = num_decoders, num_classes, dropout, pre_norm
        self.shared_decoder, self.sample_sizes = shared_decoder, sample_sizes
        self.non_parametric_queries, self.use_np_features = non_parametric_queries, use_np_features
        self.mask_dim, self.num_heads, self.num_queries = hidden_dim, num_heads, num_queries
        self.pos_enc_type = positional_encoding_type

        self.backbone = config.backbone
        self.num_levels = len(self.hlevels)
        sizes = self.backbone.PLANES[-5:]
        self.mask_features_head = conv(self.backbone.PLANES[7], self.mask_dim, kernel_size=1, stride=1, bias=True, D=3)

        if scatter_type == "mean":
            self.scatter_fn = scatter_mean
        elif scatter_type == "max":
            self.scatter_fn = lambda mask, p2s, dim: scatter_max(mask, p2s, dim=dim)[0]
        else:
            assert False, "Scatter function not known"
        assert (not use_np_features) or non_parametric_queries, "np features only with np queries"

        if non_parametric_queries:
            self.query_projection = GenericMLP(input_dim=self.mask_dim, hidden_dims=[self.mask_dim], output_dim=self.mask_dim,
                                               use_conv=True, output_use_activation=True, hidden_use_bias=True)
            if use_np_features:
                self.np_feature_projection = nn.Sequential(nn.Linear(sizes[-1], hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim))
        elif random_query_both:
            self.query_projection = GenericMLP(input_dim=2 * self.mask_dim, hidden_dims=[2 * self.mask_dim], output_dim=2 * self.mask_dim,
                                               use_conv=True, output_use_activation=True, hidden_use_bias=True)
        else:
            self.query_feat = nn.Embedding(num_queries, hidden_dim)
            self.query_pos = nn.Embedding(num_queries, hidden_dim)
        if use_level_embed:
            self.level_embed = nn.Embedding(self.num_levels, hidden_dim)

        self.mask_embed_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim))
        self.class_embed_head = nn.Linear(hidden_dim, self.num_classes)

        if positional_encoding_type == "fourier":
            self.pos_enc = PositionEmbeddingCoordsSine(pos_type="fourier", d_pos=self.mask_dim, gauss_scale=gauss_scale,
                                                       normalize=normalize_pos_enc)
        elif positional_encoding_type == "sine":
            self.pos_enc = PositionEmbeddingCoordsSine(pos_type="sine", d_pos=self.mask_dim, normalize=normalize_pos_enc)
        else:
            assert False, "pos enc type not known"

        self.pooling = MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)

        self.masked_transformer_decoder = nn.ModuleList()
        self.cross_attention, self.self_attention = nn.ModuleList(), nn.ModuleList()
        self.ffn_attention, self.lin_squeeze = nn.ModuleList(), nn.ModuleList()
        for _ in range(num_decoders if not shared_decoder else 1):
            ca, sa, ffn, sq = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
            for hlevel in self.hlevels:
                ca.append(CrossAttentionLayer(d_model=self.mask_dim, nhead=num_heads, dropout=dropout, normalize_before=pre_norm))
                sq.append(nn.Linear(sizes[hlevel], self.mask_dim))
                sa.append(SelfAttentionLayer(d_model=self.mask_dim, nhead=num_heads, dropout=dropout, normalize_before=pre_norm))
                ffn.append(FFNLayer(d_model=self.mask_dim, dim_feedforward=dim_feedforward, dropout=dropout, normalize_before=pre_norm))
            self.cross_attention.append(ca)
            self.self_attention.append(sa)
            self.ffn_attention.append(ffn)
            self.lin_squeeze.append(sq)
        self.decoder_norm = nn.LayerNorm(hidden_dim)

    # ------------------------------------------------------------------------------------------------
    def get_pos_encs(self, coords):
        """Fourier encodings of the pooled raw coordinates, per level and scene (models/mask3d.py:183-198).  The reference
        encodes all five levels and reads only those in `hlevels`; the unused ones (the full-resolution level: [N, 128] fp32
        per scene) are skipped here."""
        out = []
        for i, level in enumerate(coords):
            if i not in self.hlevels:
                out.append([None])
                continue
            per_scene = []
            for xyz in level.decomposed_features:
                lo, hi = xyz.min(dim=0)[0][None, ...], xyz.max(dim=0)[0][None, ...]
                with torch.autocast(device_type=xyz.device.type, enabled=False):
                    enc = self.pos_enc(xyz[None, ...].float(), input_range=[lo, hi])
                per_scene.append(enc.squeeze(0).permute((1, 0)))
            out.append([per_scene])
        return out

    def forward(self, x, point2segment=None, raw_coordinates=None, is_eval=False):
        if self.train_on_segments and point2segment is not None and type(self).segment_attention_core is None:
            _cuda_prepare_segment_attention(x, point2segment, len(self.sample_sizes) - 1 - min(self.hlevels))  # pooling steps of the coarsest level used
        pcd_features, aux = self.backbone(x)
        n_scenes = len(x.decomposed_coordinates)

        with torch.no_grad():
            coordinates = me.SparseTensor(features=raw_coordinates, coordinate_manager=aux[-1].coordinate_manager,
                                          coordinate_map_key=aux[-1].coordinate_map_key, device=aux[-1].device)
            coords = [coordinates]
            for _ in range(len(aux) - 1):
                coords.append(self.pooling(coords[-1]))
            coords.reverse()

        pos_encodings_pcd = self.get_pos_encs(coords)
        mask_features = self.mask_features_head(pcd_features)
        mask_segments = None
        if self.train_on_segments:
            mask_segments = [self.scatter_fn(feat, point2segment[i], dim=0)
                             for i, feat in enumerate(mask_features.decomposed_features)]

        sampled_coords = None
        if self.non_parametric_queries:
            voxel_xyz = x.decomposed_coordinates
            raw_xyz = coordinates.decomposed_features
            fps_idx = [furthest_point_sample(voxel_xyz[i][None, ...].float(), self.num_queries).squeeze(0).long()
                       for i in range(n_scenes)]
            sampled_coords = torch.stack([raw_xyz[i][fps_idx[i], :] for i in range(n_scenes)])
            mins = torch.stack([r.min(dim=0)[0] for r in raw_xyz])
            maxs = torch.stack([r.max(dim=0)[0] for r in raw_xyz])
            query_pos = self.query_projection(self.pos_enc(sampled_coords.float(), input_range=[mins, maxs]))
            if not self.use_np_features:
                queries = torch.zeros_like(query_pos).permute((0, 2, 1))
            else:
                queries = self.np_feature_projection(
                    torch.stack([pcd_features.decomposed_features[i][fps_idx[i], :] for i in range(n_scenes)]))
            query_pos = query_pos.permute((2, 0, 1))
        elif self.random_queries:
            query_pos = torch.rand(n_scenes, self.mask_dim, self.num_queries, device=x.device) - 0.5
            queries = torch.zeros_like(query_pos).permute((0, 2, 1))
            query_pos = query_pos.permute((2, 0, 1))
        elif self.random_query_both:
            shape = (n_scenes, 2 * self.mask_dim, self.num_queries)
            both = torch.randn(*shape, device=x.device) if self.random_normal else torch.rand(*shape, device=x.device) - 0.5
            queries = both[:, :self.mask_dim, :].permute((0, 2, 1))
            query_pos = both[:, self.mask_dim:, :].permute((2, 0, 1))
        else:
            queries = self.query_feat.weight.unsqueeze(0).repeat(n_scenes, 1, 1)
            query_pos = self.query_pos.weight.unsqueeze(1).repeat(1, n_scenes, 1)

        predictions_class, predictions_mask = [], []
        p2s = point2segment if self.train_on_segments else None
        segs = mask_segments if self.train_on_segments else None
        for decoder_counter in range(self.num_decoders):
            if self.shared_decoder:
                decoder_counter = 0
            for i, hlevel in enumerate(self.hlevels):
                output_class, outputs_mask, attn_mask = self.mask_module(
                    queries, mask_features, segs, len(aux) - hlevel - 1, ret_attn_mask=True, point2segment=p2s, coords=coords)
                decomposed_aux = aux[hlevel].decomposed_features
                decomposed_attn = attn_mask.decomposed_features
                sizes = [pcd.shape[0] for pcd in decomposed_aux]
                if min(sizes) == 1:
                    raise RuntimeError("only a single point gives nans in cross-attention")
                k_sample = max(sizes)
                if not (self.max_sample_size or is_eval):
                    k_sample = min(k_sample, self.sample_sizes[hlevel])

                rand_idx, mask_idx = [], []
                for n_k in sizes:
                    if n_k <= k_sample:  # take everything, pad with row 0 and mask the padding
                        idx = torch.zeros(k_sample, dtype=torch.long, device=queries.device)
                        midx = torch.ones(k_sample, dtype=torch.bool, device=queries.device)
                        idx[:n_k] = torch.arange(n_k, device=queries.device)
                        midx[:n_k] = False
                    else:  # random subset, nothing to mask
                        idx = torch.randperm(n_k, device=queries.device)[:k_sample]
                        midx = torch.zeros(k_sample, dtype=torch.bool, device=queries.device)
                    rand_idx.append(idx)
                    mask_idx.append(midx)

                batched_aux = torch.stack([decomposed_aux[k][rand_idx[k], :] for k in range(n_scenes)])
                batched_attn = torch.stack([decomposed_attn[k][rand_idx[k], :] for k in range(n_scenes)])
                batched_pos_enc = torch.stack([pos_encodings_pcd[hlevel][0][k][rand_idx[k], :] for k in range(n_scenes)])
                # queries that would attend to nothing attend to everything
                batched_attn.permute((0, 2, 1))[batched_attn.sum(1) == rand_idx[0].shape[0]] = False
                batched_attn = torch.logical_or(batched_attn, torch.stack(mask_idx)[..., None])

                src_pcd = self.lin_squeeze[decoder_counter][i](batched_aux.permute((1, 0, 2)))
                if self.use_level_embed:
                    src_pcd += self.level_embed.weight[i]
                # the reference expands the mask to [B*h, Q, K] here (:358); the attention core reads [B, K, Q] in place
                output = self.cross_attention[decoder_counter][i](
                    queries.permute((1, 0, 2)), src_pcd, memory_mask=HeadSharedMask(batched_attn),
                    memory_key_padding_mask=None, pos=batched_pos_enc.permute((1, 0, 2)), query_pos=query_pos)
                output = self.self_attention[decoder_counter][i](output, tgt_mask=None, tgt_key_padding_mask=None, query_pos=query_pos)
                queries = self.ffn_attention[decoder_counter][i](output).permute((1, 0, 2))
                predictions_class.append(output_class)
                predictions_mask.append(outputs_mask)

        output_class, outputs_mask = self.mask_module(queries, mask_features, segs, 0, ret_attn_mask=False,
                                                      point2segment=p2s, coords=coords)
        predictions_class.append(output_class)
        predictions_mask.append(outputs_mask)
        return {
            "pred_logits": predictions_class[-1],
            "pred_masks": predictions_mask[-1],
            "aux_outputs": self._set_aux_loss(predictions_class, predictions_mask),
            "sampled_coords": sampled_coords.detach().cpu().numpy() if sampled_coords is not None else None,
            "backbone_features": pcd_features,
        }

    def mask_module(self, query_feat, mask_features, mask_segments, num_pooling_steps, ret_attn_mask=True,
                    point2segment=None, coords=None):
        query_feat = self.decoder_norm(query_feat)
        mask_embed = self.mask_embed_head(query_feat)
        outputs_class = self.class_embed_head(query_feat)

        output_masks, output_segments = [], []
        if point2segment is not None:
            for i in range(len(mask_segments)):
                output_segments.append(mask_segments[i] @ mask_embed[i].T)
        else:
            per_scene = mask_features.decomposed_features
            for i in range(int(mask_features.C[-1, 0]) + 1):
                output_masks.append(per_scene[i] @ mask_embed[i].T)
        if point2segment is not None:
            if not ret_attn_mask:
                return outputs_class, output_segments
            core = type(self).segment_attention_core
            if core is None:
                core = _cuda_segment_attention_core
            return outputs_class, output_segments, core(self, mask_features, output_segments, point2segment, num_pooling_steps)
        outputs_mask = me.SparseTensor(features=torch.cat(output_masks), coordinate_manager=mask_features.coordinate_manager,
                                       coordinate_map_key=mask_features.coordinate_map_key)
        result_masks = outputs_mask.decomposed_features
        if not ret_attn_mask:
            return outputs_class, result_masks
        attn_mask = outputs_mask
        for _ in range(num_pooling_steps):
            attn_mask = self.pooling(attn_mask.float())
        attn_mask = me.SparseTensor(features=(attn_mask.F.detach().sigmoid() < 0.5),
                                    coordinate_manager=attn_mask.coordinate_manager,
                                    coordinate_map_key=attn_mask.coordinate_map_key)
        return outputs_class, result_masks, attn_mask

    @torch.jit.unused
    def _set_aux_loss(self, outputs_class, outputs_seg_masks):
        return [{"pred_logits": a, "pred_masks": b} for a, b in zip(outputs_class[:-1], outputs_seg_masks[:-1])]
