"""CPU fp32 oracle of Far3D's per-frame inference forward (TEST INFRASTRUCTURE).

Plain `torch.nn` restatement, eval-mode only, of the reference modules on the hot
path.  Parameter / buffer names equal the
reference's `state_dict` keys so one checkpoint dict drives both this oracle and
the CUDA product modules in `far3d_b200.plugin`.

All citations are relative to /root/reference/projects/mmdet3d_plugin/ unless a
third-party package is named.  Pinned against the reference's own modules run on CPU - see oracle/__init__.py.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .msda import msda_grid_sample

# --------------------------------------------------------------------------- VoVNet
# models/backbones/vovnet.py:79-87 (V-99-eSE spec)
V99 = dict(stem=[64, 64, 128], conv_ch=[128, 160, 192, 224], out_ch=[256, 512, 768, 1024],
           layers=5, blocks=[1, 3, 9, 3])
# small spec of the same family for fast CPU tests (vovnet.py:49-57, V-19-eSE)
V19 = dict(stem=[64, 64, 128], conv_ch=[128, 160, 192, 224], out_ch=[256, 512, 768, 1024],
           layers=3, blocks=[1, 1, 1, 1])
SPECS = {'V-99-eSE': V99, 'V-19-eSE': V19,
         'V-39-eSE': dict(V99, blocks=[1, 1, 2, 2]), 'V-57-eSE': dict(V99, blocks=[1, 1, 4, 3])}


def _cbr(cin, cout, name, k, stride=1):
    """conv (bias=False) + BN + ReLU triple with the reference's slash-names, vovnet.py:124-161."""
    return [(f'{name}/conv', nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)),
            (f'{name}/norm', nn.BatchNorm2d(cout)), (f'{name}/relu', nn.ReLU(inplace=True))]


class ESE(nn.Module):
    """vovnet.py:164-185: x * relu6(fc(avgpool(x)) + 3) / 6."""

    def __init__(self, c):
        super().__init__()
        self.fc = nn.Conv2d(c, c, 1)

    def forward(self, x):
        s = self.fc(x.mean((2, 3), keepdim=True))
        return x * (F.relu6(s + 3.0) / 6.0)


class OSAModule(nn.Module):
    """vovnet.py:188-238 (non-depthwise branch)."""

    def __init__(self, cin, cmid, cout, nlayers, name, identity):
        super().__init__()
        self.identity = identity
        self.layers = nn.ModuleList()
        c = cin
        for i in range(nlayers):
            self.layers.append(nn.Sequential(OrderedDict(_cbr(c, cmid, f'{name}_{i}', 3))))
            c = cmid
        self.concat = nn.Sequential(OrderedDict(_cbr(cin + nlayers * cmid, cout, f'{name}_concat', 1)))
        self.ese = ESE(cout)

    def forward(self, x):
        feats = [x]
        y = x
        for layer in self.layers:
            y = layer(y)
            feats.append(y)
        y = self.ese(self.concat(torch.cat(feats, 1)))
        return y + x if self.identity else y


class VoVNet(nn.Module):
    """vovnet.py:276-384.  Returns the list of stage outputs named in out_features."""

    def __init__(self, spec_name='V-99-eSE', input_ch=3, out_features=('stage2', 'stage3', 'stage4', 'stage5'),
                 **_):
        super().__init__()
        sp = SPECS[spec_name]
        st = sp['stem']
        self.stem = nn.Sequential(OrderedDict(_cbr(input_ch, st[0], 'stem_1', 3, 2) + _cbr(st[0], st[1], 'stem_2', 3, 1)
                                              + _cbr(st[1], st[2], 'stem_3', 3, 2)))
        cin = [st[2]] + sp['out_ch'][:-1]
        self.stage_names = []
        for i in range(4):
            stage = nn.Sequential()
            if i > 0:   # vovnet.py:248-249
                stage.add_module('Pooling', nn.MaxPool2d(3, 2, ceil_mode=True))
            for b in range(sp['blocks'][i]):
                name = f'OSA{i + 2}_{b + 1}'
                stage.add_module(name, OSAModule(cin[i] if b == 0 else sp['out_ch'][i], sp['conv_ch'][i],
                                                 sp['out_ch'][i], sp['layers'], name, identity=b > 0))
            self.add_module(f'stage{i + 2}', stage)
            self.stage_names.append(f'stage{i + 2}')
        self._out_features = tuple(out_features)

    def forward(self, x):
        outs = []
        x = self.stem(x)
        for n in self.stage_names:
            x = getattr(self, n)(x)
            if n in self._out_features:
                outs.append(x)
        return outs


# --------------------------------------------------------------------------- FPN (mmdet 2.28.2, third-party)
class _Conv(nn.Module):
    """mmcv ConvModule without norm/activation: only a `.conv` child (keys `*.conv.weight`)."""

    def __init__(self, cin, cout, k, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2)

    def forward(self, x):
        return self.conv(x)


class FPN(nn.Module):
    """mmdet FPN.forward restated for cfg far3d.py:50-57 (start_level, add_extra_convs='on_output',
    nearest top-down).  In-repo look-alike: models/necks/cp_fpn.py:156-208."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, add_extra_convs=False,
                 relu_before_extra_convs=False, **_):
        super().__init__()
        assert add_extra_convs in ('on_output', False)
        self.start_level, self.num_outs = start_level, num_outs
        self.relu_before_extra_convs = relu_before_extra_convs
        self.nlvl = len(in_channels) - start_level
        self.lateral_convs = nn.ModuleList(_Conv(c, out_channels, 1) for c in in_channels[start_level:])
        self.fpn_convs = nn.ModuleList(_Conv(out_channels, out_channels, 3) for _ in range(self.nlvl))
        for _ in range(num_outs - self.nlvl):
            self.fpn_convs.append(_Conv(out_channels, out_channels, 3, 2))

    def forward(self, inputs):
        lat = [l(inputs[i + self.start_level]) for i, l in enumerate(self.lateral_convs)]
        for i in range(self.nlvl - 1, 0, -1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
        outs = [self.fpn_convs[i](lat[i]) for i in range(self.nlvl)]
        for i in range(self.nlvl, self.num_outs):
            src = outs[-1]
            if i > self.nlvl and self.relu_before_extra_convs:
                src = F.relu(src)
            outs.append(self.fpn_convs[i](src))
        return tuple(outs)


# --------------------------------------------------------------------------- small pieces
class MLN(nn.Module):
    """models/utils/misc.py:153-190."""

    def __init__(self, c_dim, f_dim=256, use_ln=True):
        super().__init__()
        self.use_ln = use_ln
        self.reduce = nn.Sequential(nn.Linear(c_dim, f_dim), nn.ReLU())
        self.gamma = nn.Linear(f_dim, f_dim)
        self.beta = nn.Linear(f_dim, f_dim)
        if use_ln:
            self.ln = nn.LayerNorm(f_dim, elementwise_affine=False)
        nn.init.zeros_(self.gamma.weight); nn.init.zeros_(self.beta.weight)
        nn.init.ones_(self.gamma.bias); nn.init.zeros_(self.beta.bias)

    def forward(self, x, c):
        if self.use_ln:
            x = self.ln(x)
        c = self.reduce(c)
        return self.gamma(c) * x + self.beta(c)


def _sincos(p, n, temperature=10000):
    d = torch.arange(n, dtype=torch.float32, device=p.device)
    d = temperature ** (2 * torch.div(d, 2, rounding_mode='floor') / n)
    e = p[..., None] / d
    return torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1).flatten(-2)


def pos2posemb3d(pos, num_pos_feats=128, temperature=10000):
    """models/utils/positional_encoding.py:13-25; concatenation order (y, x, z)."""
    pos = pos * (2 * math.pi)
    return torch.cat([_sincos(pos[..., 1], num_pos_feats, temperature), _sincos(pos[..., 0], num_pos_feats, temperature),
                      _sincos(pos[..., 2], num_pos_feats, temperature)], dim=-1)


def pos2posemb1d(pos, num_pos_feats=256, temperature=10000):
    """positional_encoding.py:27-36."""
    return _sincos(pos[..., 0] * (2 * math.pi), num_pos_feats, temperature)


def nerf_positional_encoding(t, n=6):
    """positional_encoding.py:38-80 with include_input=False, log_sampling=True."""
    out = []
    for f in 2.0 ** torch.linspace(0.0, n - 1, n, dtype=t.dtype, device=t.device):
        out += [torch.sin(t * f), torch.cos(t * f)]
    return torch.cat(out, dim=-1)


def inverse_sigmoid(x, eps=1e-5):
    """mmdet.models.utils.transformer.inverse_sigmoid (third-party)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def topk_gather(feat, idx):
    """misc.py:13-23."""
    shape = [1] * feat.dim()
    shape[:2] = idx.shape[:2]
    return torch.gather(feat, 1, idx.view(*shape).repeat(1, 1, *feat.shape[2:]))


def transform_reference_points(pts, pose):
    """misc.py:193-202 (reverse=False, translation=True)."""
    p = torch.cat([pts, torch.ones_like(pts[..., :1])], dim=-1)
    return (pose.unsqueeze(1) @ p.unsqueeze(-1)).squeeze(-1)[..., :3]


# --------------------------------------------------------------------------- decoder
class MultiheadAttention(nn.Module):
    """mmcv.cnn.bricks.transformer.MultiheadAttention (third-party, mmcv-full 1.6.2) in eval mode,
    batch_first=True: out = identity + out_proj(softmax(q k^T/sqrt(d)) v), q/k get pos added."""

    def __init__(self, embed_dims, num_heads, **_):
        super().__init__()
        self.embed_dims = embed_dims
        self.attn = nn.MultiheadAttention(embed_dims, num_heads)

    def forward(self, query, key, value, query_pos, key_pos):
        q = (query + query_pos).transpose(0, 1)
        k = (key + key_pos).transpose(0, 1)
        out = self.attn(q, k, value.transpose(0, 1), need_weights=False)[0].transpose(0, 1)
        return query + out


class FFN(nn.Module):
    """mmcv FFN (third-party): layers = [Seq(Linear, ReLU, Dropout), Linear, Dropout]; + identity."""

    def __init__(self, embed_dims=256, feedforward_channels=1024):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                                                  nn.Dropout(0.)),
                                    nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.))

    def forward(self, x):
        return x + self.layers(x)


class DeformableFeatureAggregationCuda(nn.Module):
    """models/utils/detr3d_transformer.py:483-569, sampling through the MSDA oracle."""

    def __init__(self, embed_dims=256, num_groups=8, num_levels=4, num_cams=6, num_pts=13, **_):
        super().__init__()
        self.embed_dims, self.num_groups, self.num_levels, self.num_cams, self.num_pts = \
            embed_dims, num_groups, num_levels, num_cams, num_pts
        self.weights_fc = nn.Linear(embed_dims, num_groups * num_levels * num_pts)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.learnable_fc = nn.Linear(embed_dims, num_pts * 3)
        self.cam_embed = nn.Sequential(nn.Linear(12, embed_dims // 2), nn.ReLU(inplace=True),
                                       nn.Linear(embed_dims // 2, embed_dims), nn.ReLU(inplace=True),
                                       nn.LayerNorm(embed_dims))

    def key_points(self, x, reference_points, pc_range):       # :523-525, :27-29
        bs, nq = reference_points.shape[:2]
        ref = reference_points * (pc_range[3:6] - pc_range[0:3]) + pc_range[0:3]
        return ref.unsqueeze(-2) + self.learnable_fc(x).reshape(bs, nq, -1, 3)

    def weights(self, x, query_pos, lidar2img):                # :535-542
        bs, nq = x.shape[:2]
        cam = self.cam_embed(lidar2img[..., :3, :].flatten(-2))
        fp = (x + query_pos).unsqueeze(2) + cam.unsqueeze(1)
        w = self.weights_fc(fp).reshape(bs, nq, -1, self.num_groups).softmax(dim=-2)
        w = w.reshape(bs, nq, self.num_cams, -1, self.num_groups).permute(0, 2, 1, 4, 3).contiguous()
        return w.flatten(end_dim=1)

    def sampling_locations(self, key_points, lidar2img, pad_hw):   # :547-555
        pts = torch.cat([key_points, torch.ones_like(key_points[..., :1])], dim=-1)
        p2d = torch.matmul(lidar2img[:, :, None, None], pts[:, None, ..., None]).squeeze(-1)
        p2d = p2d[..., :2] / torch.clamp(p2d[..., 2:3], min=1e-5)
        p2d = torch.stack([p2d[..., 0] / pad_hw[1], p2d[..., 1] / pad_hw[0]], dim=-1)
        p2d = p2d.flatten(end_dim=1)
        return p2d[:, :, None, None, :, :].repeat(1, 1, self.num_groups, self.num_levels, 1, 1)

    def forward(self, x, query_pos, feat_flatten, reference_points, spatial_flatten, level_start_index, pc_range,
                lidar2img, img_metas):
        bs, nq = reference_points.shape[:2]
        kp = self.key_points(x, reference_points, pc_range)
        w = self.weights(x, query_pos, lidar2img)
        loc = self.sampling_locations(kp, lidar2img, img_metas[0]['pad_shape'][0][:2])
        bn, s, _ = feat_flatten.shape
        out = msda_grid_sample(feat_flatten.reshape(bn, s, self.num_groups, -1), spatial_flatten, level_start_index,
                               loc, w)
        out = out.reshape(bs, self.num_cams, nq, -1).sum(1)    # :565-569
        return self.output_proj(out) + x                       # :531-532 (dropout = identity in eval)


class Detr3DTemporalDecoderLayer(nn.Module):
    """detr3d_transformer.py:192-480, operation_order (self_attn, norm, cross_attn, norm, ffn, norm).
    FFN hidden size is 1024: `feedforward_channels=2048` from the config is swallowed by **kwargs (:229-245)."""

    def __init__(self, attn_cfgs, **_):
        super().__init__()
        sa, ca = [dict(c) for c in attn_cfgs]
        sa.pop('type'); ca.pop('type')
        self.attentions = nn.ModuleList([MultiheadAttention(**sa), DeformableFeatureAggregationCuda(**ca)])
        self.ffns = nn.ModuleList([FFN(256, 1024)])
        self.norms = nn.ModuleList([nn.LayerNorm(256) for _ in range(3)])

    def forward(self, query, query_pos, feats, temp_memory, temp_pos, reference_points, spatial_flatten,
                level_start_index, pc_range, lidar2img, img_metas):
        if temp_memory is not None:                              # :379-381
            kv = torch.cat([query, temp_memory], dim=1)
            kpos = torch.cat([query_pos, temp_pos], dim=1)
        else:
            kv, kpos = query, query_pos
        query = self.norms[0](self.attentions[0](query, kv, kv, query_pos, kpos))
        query = self.norms[1](self.attentions[1](query, query_pos, feats, reference_points, spatial_flatten,
                                                 level_start_index, pc_range, lidar2img, img_metas))
        return self.norms[2](self.ffns[0](query))


class Detr3DTransformerDecoder(nn.Module):
    """detr3d_transformer.py:126-190: no reference-point refinement, stacks all layer outputs."""

    def __init__(self, num_layers, transformerlayers, **_):
        super().__init__()
        cfg = dict(transformerlayers); cfg.pop('type', None)
        self.layers = nn.ModuleList([Detr3DTemporalDecoderLayer(**cfg) for _ in range(num_layers)])

    def forward(self, query, *args):
        inter = []
        for layer in self.layers:
            query = layer(query, *args)
            inter.append(query)
        return torch.stack(inter)


class Detr3DTransformer(nn.Module):
    """detr3d_transformer.py:31-124."""

    def __init__(self, decoder, **_):
        super().__init__()
        cfg = dict(decoder); cfg.pop('type', None)
        self.decoder = Detr3DTransformerDecoder(**cfg)

    def forward(self, query, query_pos, feat_flatten, spatial_flatten, level_start_index, temp_memory, temp_pos,
                attn_masks, reference_points, pc_range, data, img_metas):
        return self.decoder(query, query_pos, feat_flatten, temp_memory, temp_pos, reference_points, spatial_flatten,
                            level_start_index, pc_range, data['lidar2img'], img_metas)


# --------------------------------------------------------------------------- box coder
def denormalize_bbox(b):
    """core/bbox/util.py:25-52 (code_size 8: no velocity branch since size(-1) == 8)."""
    rot = torch.atan2(b[..., 6:7], b[..., 7:8])
    out = [b[..., 0:1], b[..., 1:2], b[..., 2:3], b[..., 3:4].exp(), b[..., 4:5].exp(), b[..., 5:6].exp(), rot]
    if b.size(-1) > 8:
        out += [b[:, 8:9], b[:, 9:10]]
    return torch.cat(out, dim=-1)


class NMSFreeCoder:
    """core/bbox/coders/nms_free_coder.py:39-112."""

    def __init__(self, pc_range, post_center_range=None, max_num=100, score_threshold=None, num_classes=10, **_):
        self.pc_range, self.post_center_range = pc_range, post_center_range
        self.max_num, self.score_threshold, self.num_classes = max_num, score_threshold, num_classes

    def decode_single(self, cls_scores, bbox_preds):
        scores, idx = cls_scores.sigmoid().view(-1).topk(self.max_num)
        labels = idx % self.num_classes
        q = torch.div(idx, self.num_classes, rounding_mode='floor')
        boxes = denormalize_bbox(bbox_preds[q])
        r = torch.tensor(self.post_center_range, device=scores.device)
        mask = (boxes[..., :3] >= r[:3]).all(1) & (boxes[..., :3] <= r[3:]).all(1)
        if self.score_threshold:
            mask &= scores >= self.score_threshold
        return dict(bboxes=boxes[mask], scores=scores[mask], labels=labels[mask], query_index=q[mask])

    def decode(self, preds):
        cls, box = preds['all_cls_scores'][-1], preds['all_bbox_preds'][-1]
        return [self.decode_single(cls[i], box[i]) for i in range(cls.size(0))]


# --------------------------------------------------------------------------- FarHead (inference)
class FarHead(nn.Module):
    """models/dense_heads/farhead.py, eval-mode forward only (:533-693), memory bank (:446-508),
    temporal alignment (:284-313), 2D->3D query lifting (:710-827), get_bboxes (:1224-1245)."""

    def __init__(self, num_classes, in_channels=256, embed_dims=256, num_query=100, memory_len=1024,
                 topk_proposals=256, num_propagated=256, with_ego_pos=True, add_query_from_2d=False,
                 depthnet_config=None, add_multi_depth_proposal=False, multi_depth_config=None,
                 return_context_feat=False, return_bbox2d_scores=False, num_reg_fcs=2, transformer=None,
                 code_weights=None, bbox_coder=None, code_size=10, **_):
        super().__init__()
        self.num_classes, self.in_channels, self.embed_dims = num_classes, in_channels, embed_dims
        self.num_query, self.memory_len = num_query, memory_len
        self.topk_proposals, self.num_propagated = topk_proposals, num_propagated
        self.with_ego_pos, self.add_query_from_2d = with_ego_pos, add_query_from_2d
        self.depthnet_config = depthnet_config or {}
        self.add_multi_depth_proposal = add_multi_depth_proposal
        self.multi_depth_config = multi_depth_config or {}
        self.return_context_feat, self.return_bbox2d_scores = return_context_feat, return_bbox2d_scores
        self.code_size = code_size
        cw = (code_weights or [1.0] * 8 + [0.2, 0.2])[:code_size]
        tcfg = dict(transformer); tcfg.pop('type', None)
        self.transformer = Detr3DTransformer(**tcfg)
        self.code_weights = nn.Parameter(torch.tensor(cw), requires_grad=False)
        self.match_costs = nn.Parameter(torch.tensor(cw), requires_grad=False)
        ccfg = dict(bbox_coder); ccfg.pop('type', None)
        self.bbox_coder = NMSFreeCoder(**ccfg)
        self.pc_range = nn.Parameter(torch.tensor(self.bbox_coder.pc_range), requires_grad=False)
        # :228-282
        cls = []
        for _ in range(num_reg_fcs):
            cls += [nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True)]
        cls.append(nn.Linear(embed_dims, num_classes))
        reg = []
        for _ in range(num_reg_fcs):
            reg += [nn.Linear(embed_dims, embed_dims), nn.ReLU()]
        reg.append(nn.Linear(embed_dims, code_size))
        fc_cls, fc_reg = nn.Sequential(*cls), nn.Sequential(*reg)
        self.cls_branches = nn.ModuleList([fc_cls for _ in range(6)])   # 6 aliases of one module (:248-251)
        self.reg_branches = nn.ModuleList([fc_reg for _ in range(6)])
        self.reference_points = nn.Embedding(num_query, 3)
        if num_propagated > 0:
            self.pseudo_reference_points = nn.Embedding(num_propagated, 3)
        self.spatial_alignment = MLN(14, use_ln=False)
        if return_context_feat or return_bbox2d_scores:
            cin = in_channels + 1 if (return_context_feat and return_bbox2d_scores) else in_channels
            self.context_embed = nn.Sequential(nn.Linear(cin, embed_dims), nn.ReLU(), nn.Linear(embed_dims, embed_dims))
        self.query_embedding = nn.Sequential(nn.Linear(embed_dims * 3 // 2, embed_dims), nn.ReLU(),
                                             nn.Linear(embed_dims, embed_dims))
        self.time_embedding = nn.Sequential(nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims))
        if with_ego_pos:
            self.ego_pose_pe = MLN(180)
            self.ego_pose_memory = MLN(180)
        self.reset_memory()

    def init_weights(self):
        """farhead.py:432-444 + detr3d_transformer.py:49-56,517-520."""
        nn.init.uniform_(self.reference_points.weight.data, 0, 1)
        if self.num_propagated > 0:
            nn.init.uniform_(self.pseudo_reference_points.weight.data, 0, 1)
        for p in self.transformer.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.transformer.modules():
            if isinstance(m, DeformableFeatureAggregationCuda):
                nn.init.zeros_(m.weights_fc.weight); nn.init.zeros_(m.weights_fc.bias)
                nn.init.xavier_uniform_(m.output_proj.weight); nn.init.zeros_(m.output_proj.bias)
                nn.init.uniform_(m.learnable_fc.bias.data, -2., 2.)     # cfg bias=2. (far3d.py:125)
        nn.init.constant_(self.cls_branches[0][-1].bias, float(-math.log((1 - 0.01) / 0.01)))

    # ---- memory bank
    def reset_memory(self):
        self.memory_embedding = self.memory_reference_point = self.memory_timestamp = None
        self.memory_egopose = self.memory_velo = None

    def pre_update_memory(self, data):                          # :453-477
        x = data['prev_exists']
        B = x.size(0)
        pr = self.pc_range
        if self.memory_embedding is None:
            self.memory_embedding = x.new_zeros(B, self.memory_len, self.embed_dims)
            self.memory_reference_point = x.new_zeros(B, self.memory_len, 3)
            self.memory_timestamp = x.new_zeros(B, self.memory_len, 1)
            self.memory_egopose = x.new_zeros(B, self.memory_len, 4, 4)
            self.memory_velo = x.new_zeros(B, self.memory_len, 2)
        else:
            self.memory_timestamp = self.memory_timestamp + data['timestamp'].unsqueeze(-1).unsqueeze(-1)
            self.memory_egopose = data['ego_pose_inv'].unsqueeze(1) @ self.memory_egopose
            self.memory_reference_point = transform_reference_points(self.memory_reference_point, data['ego_pose_inv'])
            n = self.memory_len
            self.memory_timestamp = self.memory_timestamp[:, :n] * x.view(-1, 1, 1)
            self.memory_reference_point = self.memory_reference_point[:, :n] * x.view(-1, 1, 1)
            self.memory_embedding = self.memory_embedding[:, :n] * x.view(-1, 1, 1)
            self.memory_egopose = self.memory_egopose[:, :n] * x.view(-1, 1, 1, 1)
            self.memory_velo = self.memory_velo[:, :n] * x.view(-1, 1, 1)
        if self.num_propagated > 0:
            k = self.num_propagated
            pseudo = self.pseudo_reference_points.weight * (pr[3:6] - pr[0:3]) + pr[0:3]
            self.memory_reference_point = self.memory_reference_point.clone()
            self.memory_egopose = self.memory_egopose.clone()
            self.memory_reference_point[:, :k] = self.memory_reference_point[:, :k] + (1 - x).view(B, 1, 1) * pseudo
            self.memory_egopose[:, :k] = self.memory_egopose[:, :k] + (1 - x).view(B, 1, 1, 1) * torch.eye(4, device=x.device)

    def post_update_memory(self, data, rec_ego_pose, all_cls_scores, all_bbox_preds, outs_dec):   # :479-508
        rec_ref = all_bbox_preds[..., :3][-1]
        rec_velo = all_bbox_preds[..., -2:][-1]
        rec_memory = outs_dec[-1]
        rec_score = all_cls_scores[-1].sigmoid().topk(1, dim=-1).values[..., 0:1]
        rec_ts = torch.zeros_like(rec_score, dtype=torch.float64)
        _, idx = torch.topk(rec_score, self.topk_proposals, dim=1)
        rec_ts = topk_gather(rec_ts, idx)
        rec_ref = topk_gather(rec_ref, idx)
        rec_memory = topk_gather(rec_memory, idx)
        rec_ego_pose = topk_gather(rec_ego_pose, idx)
        rec_velo = topk_gather(rec_velo, idx)
        self.last_topk_indexes = idx
        self.memory_embedding = torch.cat([rec_memory, self.memory_embedding], dim=1)
        self.memory_timestamp = torch.cat([rec_ts, self.memory_timestamp], dim=1)
        self.memory_egopose = torch.cat([rec_ego_pose, self.memory_egopose], dim=1)
        self.memory_reference_point = torch.cat([rec_ref, self.memory_reference_point], dim=1)
        self.memory_velo = torch.cat([rec_velo, self.memory_velo], dim=1)
        self.memory_reference_point = transform_reference_points(self.memory_reference_point, data['ego_pose'])
        self.memory_timestamp = self.memory_timestamp - data['timestamp'].unsqueeze(-1).unsqueeze(-1)
        self.memory_egopose = data['ego_pose'].unsqueeze(1) @ self.memory_egopose

    def temporal_alignment(self, query_pos, tgt, reference_points):       # :284-313
        B = query_pos.size(0)
        pr = self.pc_range
        temp_ref = (self.memory_reference_point - pr[:3]) / (pr[3:6] - pr[0:3])
        temp_pos = self.query_embedding(pos2posemb3d(temp_ref))
        temp_memory = self.memory_embedding
        rec_ego_pose = torch.eye(4, device=query_pos.device).view(1, 1, 4, 4).repeat(B, query_pos.size(1), 1, 1)
        if self.with_ego_pos:
            rec_motion = torch.cat([torch.zeros_like(reference_points[..., :3]), rec_ego_pose[..., :3, :].flatten(-2)], dim=-1)
            rec_motion = nerf_positional_encoding(rec_motion)
            tgt = self.ego_pose_memory(tgt, rec_motion)
            query_pos = self.ego_pose_pe(query_pos, rec_motion)
            mem_motion = torch.cat([self.memory_velo, self.memory_timestamp, self.memory_egopose[..., :3, :].flatten(-2)],
                                   dim=-1).float()
            mem_motion = nerf_positional_encoding(mem_motion)
            temp_pos = self.ego_pose_pe(temp_pos, mem_motion)
            temp_memory = self.ego_pose_memory(temp_memory, mem_motion)
        query_pos = query_pos + self.time_embedding(pos2posemb1d(torch.zeros_like(reference_points[..., :1])))
        temp_pos = temp_pos + self.time_embedding(pos2posemb1d(self.memory_timestamp).float())
        if self.num_propagated > 0:
            k = self.num_propagated
            tgt = torch.cat([tgt, temp_memory[:, :k]], dim=1)
            query_pos = torch.cat([query_pos, temp_pos[:, :k]], dim=1)
            reference_points = torch.cat([reference_points, temp_ref[:, :k]], dim=1)
            rec_ego_pose = torch.eye(4, device=query_pos.device).view(1, 1, 4, 4).repeat(B, query_pos.shape[1] + k, 1, 1)
            temp_memory = temp_memory[:, k:]
            temp_pos = temp_pos[:, k:]
        return tgt, query_pos, reference_points, temp_memory, temp_pos, rec_ego_pose

    # ---- feature prep (:553-567)
    def flatten_features(self, mlvl_feats, data):
        intr = data['intrinsics'] / 1e3
        extr = data['extrinsics'][..., :3, :]
        mln_in = torch.cat([intr[..., 0, 0:1], intr[..., 1, 1:2], extr.flatten(-2)], dim=-1).flatten(0, 1).unsqueeze(1)
        feats, shapes = [], []
        for f in mlvl_feats:
            B, N, C, H, W = f.shape
            feats.append(self.spatial_alignment(f.reshape(B * N, C, -1).transpose(1, 2), mln_in))
            shapes.append((H, W))
        feat_flatten = torch.cat(feats, dim=1)
        spatial = torch.as_tensor(shapes, dtype=torch.long, device=feat_flatten.device)
        start = torch.cat((spatial.new_zeros((1,)), spatial.prod(1).cumsum(0)[:-1]))
        return feat_flatten, spatial, start

    def bin_depth(self, idx):                                   # :521-527 (LID, bin -> depth)
        dmin, dmax, nb = [self.depthnet_config[k] for k in ('depth_min', 'depth_max', 'num_depth_bins')]
        bs = 2 * (dmax - dmin) / (nb * (1 + nb))
        return dmin + bs / 8 * (torch.square(idx / 0.5 + 1) - 1)

    def depth_to_bin(self, d):                                  # :528-531
        dmin, dmax, nb = [self.depthnet_config[k] for k in ('depth_min', 'depth_max', 'num_depth_bins')]
        bs = 2 * (dmax - dmin) / (nb * (1 + nb))
        return (-0.5 + 0.5 * torch.sqrt(1 + 8 * (d - dmin) / bs)).type(torch.int64)

    @torch.no_grad()
    def build_query2d_proposal(self, bbox_list, pred_depth, data, bn, padHW, context2d_feat, bbox2d_scores):
        """:710-827 for the configuration far3d.py uses: depth logits in, multi-depth topk=1, B == 1."""
        B, N = bn
        pad_h, pad_w = padHW
        down = int(pad_h / pred_depth.shape[1])
        nums = [len(b) for b in bbox_list]
        if sum(nums) == 0:
            return None, None
        boxes = torch.cat(bbox_list, dim=0).float()
        h_max, w_max = pred_depth.shape[1:3]
        depths = []
        for i, bb in enumerate(bbox_list):
            if nums[i] == 0:
                continue
            dm = pred_depth[i].flatten(0, 1)
            c = (bb[:, :2] / down).round().long()
            c = c.clamp(min=0)
            c[:, 0] = c[:, 0].clamp(max=w_max - 1)
            c[:, 1] = c[:, 1].clamp(max=h_max - 1)
            flat = (c[:, 1] * (pad_w / down) + c[:, 0]).long()
            depths.append(dm[flat])
        depths = torch.cat(depths, dim=0)
        topk = self.multi_depth_config.get('topk', -1)
        use_logits = topk != -1
        if self.add_multi_depth_proposal:
            if use_logits:
                range_min_bin = self.depth_to_bin(torch.tensor([float(self.multi_depth_config.get('range_min', -1))])).item()
                tv, ti = torch.topk(depths, topk, dim=1)
                ok = ti[:, 0] >= range_min_bin
                boxes = torch.cat([boxes, boxes.repeat(topk - 1, 1)[ok.repeat(topk - 1)]], dim=0)
                extra = ti[:, 1:][ok].transpose(1, 0).flatten().unsqueeze(-1)
                depths = torch.cat([ti[:, 0:1], extra], dim=0)
                if context2d_feat is not None:
                    context2d_feat = torch.cat([context2d_feat, context2d_feat.repeat(topk - 1, 1)[ok.repeat(topk - 1)]], dim=0)
            if bbox2d_scores is not None:
                thr = torch.tensor([0.1]).to(bbox2d_scores.device)
                log_odds = torch.log(bbox2d_scores / (1 - bbox2d_scores)) - torch.log(thr / (1 - thr))
                if use_logits:
                    tv = tv / tv[:, 0:1]
                    ds = torch.cat([tv[:, 0:1], tv[:, 1:][ok].transpose(1, 0).flatten().unsqueeze(-1)], dim=0)
                    log_odds = torch.cat([log_odds, log_odds[ok].repeat(topk - 1, 1)], dim=0) * ds
                context2d_feat = torch.cat([context2d_feat, log_odds], dim=-1) if context2d_feat is not None \
                    else log_odds.repeat(1, self.in_channels)
        depths = self.bin_depth(depths)
        coords = torch.cat([boxes[:, :2], depths], dim=1)
        coords = torch.cat((coords, torch.ones_like(coords[..., :1])), -1)
        coords[..., :2] = coords[..., :2] * torch.maximum(coords[..., 2:3], torch.ones_like(coords[..., 2:3]) * 1e-5)
        img2lidars = data['lidar2img'].inverse().view(B * N, 1, 4, 4)
        i2l = torch.cat([img2lidars[k].repeat(n, 1, 1) for k, n in enumerate(nums)], dim=0)
        if self.add_multi_depth_proposal and use_logits:
            i2l = torch.cat([i2l, i2l.repeat(topk - 1, 1, 1)[ok.repeat(topk - 1)]], dim=0)
        c3 = torch.matmul(i2l, coords.unsqueeze(-1)).squeeze(-1)[..., :3]
        pr = self.pc_range
        c3 = (c3 - pr[:3]) / (pr[3:6] - pr[:3])
        assert B == 1
        return c3.unsqueeze(0), (context2d_feat.unsqueeze(0) if context2d_feat is not None else None)

    def forward(self, img_metas, outs_roi=None, **data):
        self.pre_update_memory(data)
        mlvl_feats = data['img_feats']
        B, N = mlvl_feats[0].shape[:2]
        feat_flatten, spatial, start = self.flatten_features(mlvl_feats, data)
        reference_points = self.reference_points.weight.unsqueeze(0).repeat(B, 1, 1)     # :424-427
        query_pos = self.query_embedding(pos2posemb3d(reference_points))
        ref2d = ctx = None
        npro = 0
        if self.add_query_from_2d and outs_roi is not None:
            pred_depth = outs_roi['pred_depth']
            scores = outs_roi['bbox2d_scores'] if self.return_bbox2d_scores else None
            ctx2d = None
            if self.return_context_feat:
                vi = outs_roi['valid_indices']
                ctx2d = feat_flatten[vi.repeat(1, 1, feat_flatten.shape[-1])].reshape(-1, feat_flatten.shape[-1])
            padHW = img_metas[0]['pad_shape'][0][:2]
            ref2d, ctx = self.build_query2d_proposal(outs_roi['bbox_list'], pred_depth.permute(0, 2, 3, 1), data, (B, N),
                                                     padHW, ctx2d, scores)
            if ref2d is not None:
                npro = ref2d.shape[1]
                query_pos = torch.cat([query_pos, self.query_embedding(pos2posemb3d(ref2d))], dim=1)
                reference_points = torch.cat([reference_points, ref2d], dim=1)
        tgt = torch.zeros_like(query_pos)
        if ctx is not None:
            tgt[:, -npro:, :] = self.context_embed(ctx)
        tgt, query_pos, reference_points, temp_memory, temp_pos, rec_ego_pose = \
            self.temporal_alignment(query_pos, tgt, reference_points)
        outs_dec = self.transformer(tgt, query_pos, feat_flatten, spatial, start, temp_memory, temp_pos, None,
                                    reference_points, self.pc_range, data, img_metas)
        outs_dec = torch.nan_to_num(outs_dec)
        ref_logit = inverse_sigmoid(reference_points.clone())
        cls_all, box_all = [], []
        for lvl in range(outs_dec.shape[0]):
            cls_all.append(self.cls_branches[lvl](outs_dec[lvl]))
            tmp = self.reg_branches[lvl](outs_dec[lvl]).clone()
            tmp[..., 0:3] = (tmp[..., 0:3] + ref_logit[..., 0:3]).sigmoid()
            box_all.append(tmp)
        all_cls = torch.stack(cls_all)
        all_box = torch.stack(box_all)
        pr = self.pc_range
        all_box[..., 0:3] = all_box[..., 0:3] * (pr[3:6] - pr[0:3]) + pr[0:3]
        self.post_update_memory(data, rec_ego_pose, all_cls, all_box, outs_dec)
        return dict(all_cls_scores=all_cls, all_bbox_preds=all_box, dn_mask_dict=None, reference_points2d=ref2d,
                    outs_dec=outs_dec, feat_flatten=feat_flatten)

    def get_bboxes(self, preds, img_metas=None):               # :1224-1245 (boxes returned as plain tensors)
        ret = []
        for p in self.bbox_coder.decode(preds):
            b = p['bboxes'].clone()
            b[:, 2] = b[:, 2] - b[:, 5] * 0.5
            ret.append([b, p['scores'], p['labels']])
        return ret


# --------------------------------------------------------------------------- 2D proposal head (row a6)
class _ConvBNSwish(nn.Module):
    """mmcv ConvModule(conv bias=False, BN(eps=1e-3), Swish): children `conv`, `bn`."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, 1, 1, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=0.001, momentum=0.03)

    def forward(self, x):
        x = self.bn(self.conv(x))
        return x * torch.sigmoid(x)


class DepthPredictor(nn.Module):
    """models/depth_predictor/depth_predictor.py:62-86 (single-level, conv_layer_num default 2)."""

    def __init__(self, cfg):
        super().__init__()
        d = 256
        mk = lambda a, b: nn.Sequential(nn.Conv2d(a, b, 3, padding=1), nn.GroupNorm(32, b), nn.ReLU())
        n = cfg.get('conv_layer_num', 2)
        self.depth_head = nn.Sequential(mk(cfg['hidden_dim'], d), *[mk(d, d) for _ in range(n - 1)])
        self.depth_classifier = nn.Conv2d(d, int(cfg['num_depth_bins']) + 1, 1)

    def forward(self, x):
        return self.depth_classifier(self.depth_head(x))


class YOLOXHeadCustom(nn.Module):
    """models/dense_heads/yolox_head.py: forward (:260-341) and get_bboxes (:355-489), test path."""

    def __init__(self, num_classes, in_channels, feat_channels=256, stacked_convs=2, strides=(8, 16, 32),
                 pred_with_depth=False, depthnet_config=None, reg_depth_level='p4', sample_with_score=True,
                 threshold_score=0.05, **_):
        super().__init__()
        self.num_classes, self.strides = num_classes, list(strides)
        self.threshold_score, self.sample_with_score = threshold_score, sample_with_score
        self.pred_with_depth, self.reg_depth_level = pred_with_depth, reg_depth_level
        mk = lambda: nn.Sequential(*[_ConvBNSwish(in_channels if i == 0 else feat_channels, feat_channels)
                                     for i in range(stacked_convs)])
        self.multi_level_cls_convs = nn.ModuleList(mk() for _ in strides)
        self.multi_level_reg_convs = nn.ModuleList(mk() for _ in strides)
        self.multi_level_conv_cls = nn.ModuleList(nn.Conv2d(feat_channels, num_classes, 1) for _ in strides)
        self.multi_level_conv_reg = nn.ModuleList(nn.Conv2d(feat_channels, 4, 1) for _ in strides)
        self.multi_level_conv_obj = nn.ModuleList(nn.Conv2d(feat_channels, 1, 1) for _ in strides)
        self.multi_level_conv_centers2d = nn.ModuleList(nn.Conv2d(feat_channels, 2, 1) for _ in strides)
        if pred_with_depth:
            self.depthnet = DepthPredictor(depthnet_config)

    def init_weights(self):                                     # :232-236
        b = float(-math.log((1 - 0.01) / 0.01))
        for c, o in zip(self.multi_level_conv_cls, self.multi_level_conv_obj):
            c.bias.data.fill_(b); o.bias.data.fill_(b)

    def forward(self, locations=None, **data):
        feats = data['img_feats']
        cls, box, obj, ctr = [], [], [], []
        for i, f in enumerate(feats):
            x = f.flatten(0, 1) if f.dim() == 5 else f
            cf = self.multi_level_cls_convs[i](x)
            rf = self.multi_level_reg_convs[i](x)
            cls.append(self.multi_level_conv_cls[i](cf)); box.append(self.multi_level_conv_reg[i](rf))
            obj.append(self.multi_level_conv_obj[i](rf)); ctr.append(self.multi_level_conv_centers2d[i](rf))
        out = dict(enc_cls_scores=cls, enc_bbox_preds=box, pred_centers2d_offset=ctr, objectnesses=obj, topk_indexes=None)
        if self.pred_with_depth:
            ridx = ['p3', 'p4', 'p5'].index(self.reg_depth_level)
            logit = self.depthnet(feats[ridx].flatten(0, 1))
            out.update(depth_logit=logit, pred_depth=logit.softmax(dim=1))
        return out

    def priors(self, sizes, device):
        """mmdet MlvlPointGenerator(strides, offset=0).grid_priors(with_stride=True) (third-party)."""
        res = []
        for (h, w), s in zip(sizes, self.strides):
            xs = torch.arange(0, w, device=device, dtype=torch.float32) * s
            ys = torch.arange(0, h, device=device, dtype=torch.float32) * s
            xx = xs.repeat(h); yy = ys.view(-1, 1).repeat(1, w).view(-1)
            res.append(torch.stack([xx, yy, xx.new_full(xx.shape, s), xx.new_full(xx.shape, s)], dim=-1))
        return res

    def get_bboxes(self, preds):
        cls, box, obj = preds['enc_cls_scores'], preds['enc_bbox_preds'], preds['objectnesses']
        n = cls[0].shape[0]
        pri = torch.cat(self.priors([c.shape[2:] for c in cls], cls[0].device))
        sw = []
        for i in range(len(obj)):
            w = obj[i].sigmoid() * cls[i].topk(1, dim=1).values.sigmoid()
            wn = F.max_pool2d(w, (3, 3), stride=1, padding=1).permute(0, 2, 3, 1).reshape(n, -1, 1)
            w_ = w.permute(0, 2, 3, 1).reshape(n, -1, 1)
            sw.append(w_ * (w_ == wn).float())
        score = torch.cat(sw, dim=1)
        valid = score > self.threshold_score
        bp = torch.cat([b.permute(0, 2, 3, 1).reshape(n, -1, 4) for b in box], dim=1)
        xy = bp[..., :2] * pri[:, 2:] + pri[:, :2]
        wh = bp[..., 2:].exp() * pri[:, 2:]
        boxes = torch.cat([xy - wh / 2, xy + wh / 2], dim=-1)
        res = []
        for i in range(n):
            b = boxes[i][valid[i].repeat(1, 4)].reshape(-1, 4)
            res.append(torch.cat([(b[:, :2] + b[:, 2:]) / 2, b[:, 2:] - b[:, :2]], dim=-1))
        return dict(bbox_list=res, bbox2d_scores=score[valid].reshape(-1, 1), valid_indices=valid)


# --------------------------------------------------------------------------- detector
class Far3D(nn.Module):
    """models/detectors/far3d.py: extract_img_feat (:64-99), simple_test_pts (:244-266), simple_test (:268-277)."""

    def __init__(self, img_backbone, img_neck, pts_bbox_head, img_roi_head=None, position_level=(0,), stride=(16,),
                 **_):
        super().__init__()
        def strip(c):
            c = dict(c); c.pop('type', None); c.pop('train_cfg', None); c.pop('test_cfg', None); return c
        self.img_backbone = VoVNet(**strip(img_backbone)) if not isinstance(img_backbone, nn.Module) else img_backbone
        self.img_neck = FPN(**strip(img_neck))
        self.pts_bbox_head = FarHead(**strip(pts_bbox_head))
        self.img_roi_head = YOLOXHeadCustom(**strip(img_roi_head)) if img_roi_head is not None else None
        self.position_level, self.stride = list(position_level), list(stride)
        self.prev_scene_token = None

    def extract_img_feat(self, img):
        B = img.size(0)
        if img.dim() == 5:
            img = img.flatten(0, 1)
        feats = self.img_neck(self.img_backbone(img))
        return [feats[i].view(B, feats[i].size(0) // B, *feats[i].shape[1:]) for i in self.position_level]

    @torch.no_grad()
    def simple_test(self, img_metas, inject_roi=None, **data):
        data['img_feats'] = self.extract_img_feat(data['img'])
        outs_roi = None
        if self.img_roi_head is not None:
            outs_roi = self.img_roi_head(None, **data)
            outs_roi.update(self.img_roi_head.get_bboxes(outs_roi))
        if inject_roi is not None:
            outs_roi = dict(outs_roi or {}, **inject_roi)
        if img_metas[0]['scene_token'] != self.prev_scene_token:
            self.prev_scene_token = img_metas[0]['scene_token']
            data['prev_exists'] = data['img'].new_zeros(1)
            self.pts_bbox_head.reset_memory()
        else:
            data['prev_exists'] = data['img'].new_ones(1)
        outs = self.pts_bbox_head(img_metas, outs_roi, **data)
        boxes = self.pts_bbox_head.get_bboxes(outs, img_metas)
        return [dict(pts_bbox=dict(boxes_3d=b, scores_3d=s, labels_3d=l)) for b, s, l in boxes], outs
