"""Sparse 3D-query decoder on the sm_100a kernels; same registry names, constructor and forward signatures and
state_dict keys as the reference's projects/mmdet3d_plugin/models/utils/detr3d_transformer.py (inference only).

torch.nn modules are used as parameter containers (so checkpoints load by the reference's keys); all arithmetic
runs through far3d_b200.ops -> libfar3d_sm100.so.  Dropout layers of the reference are identities at inference."""
import copy

import torch
import torch.nn as nn

from .. import ops
from ..compat import (ATTENTION, FEEDFORWARD_NETWORK, TRANSFORMER, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE,
                      build_from_cfg)


def _inference_only(m):
    if m.training:
        raise RuntimeError(f'{type(m).__name__}: far3d_b200 implements the inference forward only; call .eval()')


@ATTENTION.register_module()
class MultiheadAttention(nn.Module):
    """mmcv.cnn.bricks.transformer.MultiheadAttention (config far3d.py:112-116): in/out projections as GEMMs,
    fused softmax(QK^T)V kernel; `query + query_pos` / `key + key_pos` are folded into the GEMM operand load."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=None, init_cfg=None,
                 batch_first=False, dropout=None, **kwargs):
        super().__init__()
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads)     # parameter container (keys attn.in_proj_*, attn.out_proj.*)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, query_is_key_prefix=False, **kwargs):
        _inference_only(self)
        assert self.batch_first, 'far3d_b200 MultiheadAttention expects batch_first=True (far3d.py:111)'
        assert attn_mask is None and key_padding_mask is None, 'masks are training-only in Far3D'
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        E = self.embed_dims
        W, b = self.attn.in_proj_weight, self.attn.in_proj_bias
        Nq = query.shape[1]
        if query_is_key_prefix and query.shape[0] == 1 and key_pos is not None and query_pos is not None and key.shape[1] >= Nq:
            # the decoder's self-attention: key = cat(query, memory), key_pos = cat(query_pos, memory_pos), so the query rows
            # of the Q projection are the first Nq rows of the K projection's input - one GEMM with [W_q; W_k] gives both
            qk = ops.linear(key, W[:2 * E], b[:2 * E], x_add=key_pos)              # [1, Nk, 2E]
            q, k = qk[:, :Nq, :E], qk[:, :, E:]
        else:
            q = ops.linear(query, W[:E], b[:E], x_add=query_pos)
            k = ops.linear(key, W[E:2 * E], b[E:2 * E], x_add=key_pos)
        v = ops.linear(value, W[2 * E:], b[2 * E:])
        o = ops.mha(q, k, v, self.num_heads)
        return ops.linear(o, self.attn.out_proj.weight, self.attn.out_proj.bias, residual=identity)


@FEEDFORWARD_NETWORK.register_module()
class FFN(nn.Module):
    """mmcv FFN: Linear-ReLU-Linear + identity (ffn_drop = 0)."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=None, ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        assert num_fcs == 2
        self.embed_dims, self.add_identity = embed_dims, add_identity
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                                                  nn.Dropout(ffn_drop)),
                                    nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))

    def forward(self, x, identity=None):
        _inference_only(self)
        h = ops.linear(x, self.layers[0][0].weight, self.layers[0][0].bias, act=1)
        res = (x if identity is None else identity) if self.add_identity else None
        return ops.linear(h, self.layers[1].weight, self.layers[1].bias, residual=res)


@ATTENTION.register_module()
class DeformableFeatureAggregationCuda(nn.Module):
    """Perspective-aware aggregation, detr3d_transformer.py:483-569, on the fused kernel far3d_deform_agg_fwd.

    Differences from the reference's op sequence (same results within fp32 rounding):
      * weights_fc is linear, so logits = weights_fc(x + pos) + W.cam_embed are formed from a [Nq,416] and a [N,416]
        GEMM instead of a [Nq*N,416] one (:539-540); the softmax kernel adds them on the fly;
      * projection, bounds test, 4-level x 13-point bilinear gather, weighting and the sum over cameras are one kernel
        (no 21 MB sampling-location tensor, no per-camera outputs).
    """

    def __init__(self, embed_dims=256, num_groups=8, num_levels=4, num_cams=6, dropout=0.1, num_pts=13, im2col_step=64,
                 batch_first=True, bias=1.):
        super().__init__()
        self.embed_dims, self.num_groups, self.num_levels = embed_dims, num_groups, num_levels
        self.group_dims = embed_dims // num_groups
        self.num_cams, self.num_pts, self.im2col_step, self.bias = num_cams, num_pts, im2col_step, bias
        self.weights_fc = nn.Linear(embed_dims, num_groups * num_levels * num_pts)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.learnable_fc = nn.Linear(embed_dims, num_pts * 3)
        self.cam_embed = nn.Sequential(nn.Linear(12, embed_dims // 2), nn.ReLU(inplace=True),
                                       nn.Linear(embed_dims // 2, embed_dims), nn.ReLU(inplace=True),
                                       nn.LayerNorm(embed_dims))
        self.drop = nn.Dropout(dropout)
        self.feat_dtype = torch.float32

    def init_weight(self):          # detr3d_transformer.py:517-520
        nn.init.zeros_(self.weights_fc.weight); nn.init.zeros_(self.weights_fc.bias)
        nn.init.xavier_uniform_(self.output_proj.weight); nn.init.zeros_(self.output_proj.bias)
        nn.init.uniform_(self.learnable_fc.bias.data, -self.bias, self.bias)

    # layer-invariant operands the decoder computed once for all its layers (Detr3DTransformerDecoder.forward): None outside it
    _shared = None

    def key_points(self, instance_feature, reference_points, pc_range):
        bs, nq = reference_points.shape[:2]
        sh = self._shared
        if sh is not None and sh.get('ref_rep') is not None:
            ref_rep = sh['ref_rep']
        else:
            ref_rep = self.reference_rows(reference_points, pc_range, self.num_pts)
        kp = ops.linear(instance_feature, self.learnable_fc.weight, self.learnable_fc.bias, residual=ref_rep)
        return kp.view(bs, nq, self.num_pts, 3)

    @staticmethod
    def reference_rows(reference_points, pc_range, num_pts):
        ref = reference_points * (pc_range[3:6] - pc_range[0:3]) + pc_range[0:3]          # glue: Nq x 3
        return ref.repeat(1, 1, num_pts)                                                   # [bs, nq, P*3]

    def cam_layer_tensors(self):
        """(w0, b0, w1, b1, ln weight, ln bias, weights_fc.weight) for ops.cam_logits"""
        ce = self.cam_embed
        return (ce[0].weight, ce[0].bias, ce[2].weight, ce[2].bias, ce[4].weight, ce[4].bias, self.weights_fc.weight)

    def _weight_logits(self, instance_feature, anchor_embed, lidar2img_mat):
        """logits[b,q,n,:] = weights_fc(feature + anchor_embed + cam_embed(lidar2img)) = wq[b,q,:] + wc[b,n,:] (weights_fc is linear)"""
        bs = instance_feature.shape[0]
        wq = ops.linear(instance_feature, self.weights_fc.weight, self.weights_fc.bias, x_add=anchor_embed)
        sh = self._shared
        if sh is not None and sh.get('wc') is not None:
            wc = sh['wc']                                                                  # [bs, N, J], from ops.cam_logits
        else:
            cam_in = lidar2img_mat[..., :3, :].flatten(-2).contiguous()                    # [bs, N, 12]
            h = ops.linear(cam_in, self.cam_embed[0].weight, self.cam_embed[0].bias, act=1)
            h = ops.linear(h, self.cam_embed[2].weight, self.cam_embed[2].bias, act=1)
            cam = ops.layernorm(h, self.cam_embed[4].weight, self.cam_embed[4].bias, self.cam_embed[4].eps)
            wc = ops.linear(cam, self.weights_fc.weight, None)
        return wq.view(bs, -1, wq.shape[-1]), wc.view(bs, -1, wc.shape[-1])

    def _get_weights(self, instance_feature, anchor_embed, lidar2img_mat):
        wq, wc = self._weight_logits(instance_feature, anchor_embed, lidar2img_mat)
        return ops.dfa_weights_softmax(wq, wc, self.num_groups)

    # two-kernel form of softmax + aggregation (far3d_dfa_prepare -> far3d_deform_agg_gather): same result bit for bit
    prepared = True

    def forward(self, instance_feature, query_pos, feat_flatten, reference_points, spatial_flatten, level_start_index,
                pc_range, lidar2img_mat, img_metas):
        _inference_only(self)
        key_points = self.key_points(instance_feature, reference_points, pc_range)
        pad_h, pad_w = img_metas[0]['pad_shape'][0][:2]
        shapes, starts = _host_levels(spatial_flatten, level_start_index)
        l2i = lidar2img_mat.contiguous()
        if self.prepared and ops.dfa_prepare_supported(l2i.shape[1], self.num_groups, len(shapes), self.num_pts, feat_flatten.shape[-1]):
            wq, wc = self._weight_logits(instance_feature, query_pos, lidar2img_mat)
            feats = ops.deform_agg_prepared(feat_flatten, shapes, starts, key_points, l2i, wq, wc, pad_h, pad_w, self.num_groups)
        else:
            weights = self._get_weights(instance_feature, query_pos, lidar2img_mat)
            feats = ops.deform_agg(feat_flatten, shapes, starts, key_points, l2i, weights, pad_h, pad_w, self.num_groups)
        return ops.linear(feats, self.output_proj.weight, self.output_proj.bias, residual=instance_feature)


_LEVEL_CACHE = {}


def _host_levels(spatial_flatten, level_start_index):
    """(H,W) per level and start indices as host tuples.  Accepts host sequences (no sync) or the reference's device
    int64 tensors (one D2H copy, cached per tensor storage)."""
    if not torch.is_tensor(spatial_flatten):
        return tuple(map(tuple, spatial_flatten)), tuple(level_start_index)
    key = (spatial_flatten.data_ptr(), level_start_index.data_ptr(), spatial_flatten._version)
    hit = _LEVEL_CACHE.get(key)
    if hit is None:
        hit = (tuple(map(tuple, spatial_flatten.tolist())), tuple(level_start_index.tolist()))
        _LEVEL_CACHE.clear()
        _LEVEL_CACHE[key] = hit
    return hit


@TRANSFORMER_LAYER.register_module()
class Detr3DTemporalDecoderLayer(nn.Module):
    """detr3d_transformer.py:192-480.  The FFN keeps the reference's effective size 256->1024->256:
    `feedforward_channels` / `ffn_dropout` given in far3d.py:127-128 fall into **kwargs there too (:229-245)."""

    def __init__(self, attn_cfgs=None,
                 ffn_cfgs=dict(type='FFN', embed_dims=256, feedforward_channels=1024, num_fcs=2, ffn_drop=0.,
                               act_cfg=dict(type='ReLU', inplace=True)),
                 operation_order=None, norm_cfg=dict(type='LN'), init_cfg=None, batch_first=False, with_cp=True,
                 **kwargs):
        super().__init__()
        assert set(operation_order) <= {'self_attn', 'norm', 'ffn', 'cross_attn'}
        self.batch_first, self.operation_order = batch_first, tuple(operation_order)
        self.pre_norm = operation_order[0] == 'norm'
        num_attn = operation_order.count('self_attn') + operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
        assert num_attn == len(attn_cfgs)
        self.num_attn = num_attn
        self.attentions = nn.ModuleList()
        idx = 0
        for op in operation_order:
            if op in ('self_attn', 'cross_attn'):
                cfg = dict(attn_cfgs[idx])
                if 'batch_first' in cfg:
                    assert cfg['batch_first'] == batch_first
                else:
                    cfg['batch_first'] = batch_first
                att = build_from_cfg(cfg, ATTENTION)
                att.operation_name = op
                self.attentions.append(att)
                idx += 1
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = nn.ModuleList()
        for _ in range(operation_order.count('ffn')):
            cfg = dict(ffn_cfgs)
            cfg.setdefault('embed_dims', self.embed_dims)
            self.ffns.append(build_from_cfg(cfg, FEEDFORWARD_NETWORK))
        self.norms = nn.ModuleList(nn.LayerNorm(self.embed_dims) for _ in range(operation_order.count('norm')))
        self.use_checkpoint = with_cp

    _kpos = None          # cat(query_pos, temp_pos), the same for every layer: set by Detr3DTransformerDecoder.forward
    fuse_qk = True        # self-attention: Q and K projections as one GEMM (the query rows are a prefix of the key rows)

    def forward(self, query, query_pos, mlvl_feats, temp_memory, temp_pos, reference_points, spatial_flatten,
                level_start_index, pc_range, lidar2img, img_metas, attn_masks=None, query_key_padding_mask=None,
                key_padding_mask=None):
        _inference_only(self)
        ni = ai = fi = 0
        identity = query
        for op in self.operation_order:
            if op == 'self_attn':
                if temp_memory is not None:                       # :379-381 (glue: two small concats)
                    kv = torch.cat([query, temp_memory], dim=1)
                    kpos = self._kpos if self._kpos is not None else torch.cat([query_pos, temp_pos], dim=1)
                else:
                    kv, kpos = query, query_pos
                query = self.attentions[ai](query, kv, kv, identity if self.pre_norm else None, query_pos=query_pos,
                                            key_pos=kpos, query_is_key_prefix=self.fuse_qk)
                ai += 1
                identity = query
            elif op == 'norm':
                n = self.norms[ni]
                query = ops.layernorm(query, n.weight, n.bias, n.eps)
                ni += 1
            elif op == 'cross_attn':
                query = self.attentions[ai](query, query_pos, mlvl_feats, reference_points, spatial_flatten,
                                            level_start_index, pc_range, lidar2img, img_metas)
                ai += 1
                identity = query
            elif op == 'ffn':
                query = self.ffns[fi](query, identity if self.pre_norm else None)
                fi += 1
        return query


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class Detr3DTransformerDecoder(nn.Module):
    """detr3d_transformer.py:126-190 (mmcv TransformerLayerSequence): stacks every layer's output."""

    def __init__(self, embed_dims=None, transformerlayers=None, num_layers=None, init_cfg=None, **kwargs):
        super().__init__()
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        assert len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = nn.ModuleList(build_from_cfg(c, TRANSFORMER_LAYER) for c in transformerlayers)
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm

    def forward(self, query, query_pos, mlvl_feats, temp_memory, temp_pos, reference_points, spatial_flatten,
                level_start_index, pc_range, lidar2img, img_metas, attn_masks=None):
        inter = []
        self._share_layer_invariants(query_pos, temp_memory, temp_pos, reference_points, pc_range, lidar2img)
        try:
            for layer in self.layers:
                query = layer(query, query_pos, mlvl_feats, temp_memory, temp_pos, reference_points, spatial_flatten,
                              level_start_index, pc_range, lidar2img, img_metas, attn_masks)
                inter.append(query)
        finally:
            for layer in self.layers:
                layer._kpos = None
                for att in layer.attentions:
                    if isinstance(att, DeformableFeatureAggregationCuda):
                        att._shared = None
        return torch.stack(inter)

    hoist = True

    def _share_layer_invariants(self, query_pos, temp_memory, temp_pos, reference_points, pc_range, lidar2img):
        """What every layer would recompute from the same inputs, once: the key position rows cat(query_pos, temp_pos); the
        metric reference rows the key-point offsets are added to; and the camera side of the aggregation logits of ALL layers
        (far3d_cam_logits: one launch instead of five 7-row launches per layer)."""
        if not self.hoist:
            return
        kpos = torch.cat([query_pos, temp_pos], dim=1) if temp_memory is not None else None
        dfas = [att for layer in self.layers for att in layer.attentions if isinstance(att, DeformableFeatureAggregationCuda)]
        wc = None
        same = dfas and all(d.num_pts == dfas[0].num_pts and d.embed_dims == dfas[0].embed_dims and
                            d.cam_embed[4].eps == 1e-5 and d.weights_fc.weight.shape == dfas[0].weights_fc.weight.shape for d in dfas)
        if same and len(dfas) <= 8 and lidar2img.is_cuda and lidar2img.dtype == torch.float32:
            wc = ops.cam_logits(lidar2img.contiguous(), [d.cam_layer_tensors() for d in dfas])
        ref_rep = DeformableFeatureAggregationCuda.reference_rows(reference_points, pc_range, dfas[0].num_pts) if same else None
        for layer in self.layers:
            layer._kpos = kpos
        for i, d in enumerate(dfas):
            d._shared = dict(wc=wc[i] if wc is not None else None, ref_rep=ref_rep)


@TRANSFORMER.register_module()
class Detr3DTransformer(nn.Module):
    """detr3d_transformer.py:31-124."""

    def __init__(self, decoder=None, init_cfg=None, **kwargs):
        super().__init__()
        self.decoder = build_from_cfg(decoder, TRANSFORMER_LAYER_SEQUENCE)

    def init_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if hasattr(m, 'init_weight'):
                m.init_weight()

    def forward(self, query, query_pos, feat_flatten, spatial_flatten, level_start_index, temp_memory, temp_pos,
                attn_masks, reference_points, pc_range, data, img_metas):
        return self.decoder(query=query, query_pos=query_pos, mlvl_feats=feat_flatten, temp_memory=temp_memory,
                            temp_pos=temp_pos, reference_points=reference_points, spatial_flatten=spatial_flatten,
                            level_start_index=level_start_index, pc_range=pc_range, lidar2img=data['lidar2img'],
                            img_metas=img_metas, attn_masks=attn_masks)
