"""Stand-ins for the THIRD-PARTY packages the reference imports (mmcv-full 1.6.2, mmdet 2.28.2, mmdet3d 1.0.0rc6 - none
installable offline), so that the reference's OWN modules under /root/reference/projects/mmdet3d_plugin can be imported
and executed on CPU, unmodified, to produce golden vectors (tests/golden/make_ref_golden.py).

Test infrastructure only: used by make_ref_golden.py in the build container; nothing under far3d_b200/ imports it and it
does not travel into any product path.  Every class below restates the published behaviour of the named third-party
symbol at the pinned version, reduced to what the inference path touches (training-only pieces raise or are inert):

  mmcv.runner            BaseModule, force_fp32, auto_fp16 (no-ops while fp16_enabled is False, as in mmcv)
  mmcv.utils             Registry, build_from_cfg, ConfigDict, deprecated_api_warning
  mmcv.cnn               ConvModule (conv -> 'bn' -> 'activate'), Swish, build_norm_layer, xavier_init, constant_init,
                         bias_init_with_prob, Linear, Scale
  mmcv.cnn.bricks.transformer   MultiheadAttention (wrapper over nn.MultiheadAttention with query_pos / key_pos / identity),
                         FFN, TransformerLayerSequence, build_* helpers
  mmcv.ops.multi_scale_deform_attn   MultiScaleDeformableAttnFunction: the package's own pure-PyTorch statement of the op,
                         `multi_scale_deformable_attn_pytorch` (per level F.grid_sample(bilinear, zeros, align_corners=False))
  mmdet                  FPN, MlvlPointGenerator, inverse_sigmoid, bbox_xyxy_to_cxcywh, multi_apply, head base classes
  mmdet3d                MVXTwoStageDetector (only the sub-module building), bbox3d2result
"""
import copy
import math
import sys
import types
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ mmcv.utils
class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v

    def __setattr__(self, k, v):
        self[k] = v


def to_config(x):
    """mmcv.Config turns every nested dict of a config file into a ConfigDict (attribute access)."""
    if isinstance(x, dict):
        return ConfigDict({k: to_config(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return type(x)(to_config(v) for v in x)
    return x


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop('type')
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f'{t} is not in the {registry.name} registry')
    return cls(**args)


class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name, self.module_dict = name, {}

    def get(self, key):
        return self.module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        return _reg(module) if module is not None else _reg

    def build(self, cfg, default_args=None, **kw):
        return build_from_cfg(cfg, self, default_args)


def deprecated_api_warning(name_dict, cls_name=None):
    return lambda f: f


# ------------------------------------------------------------------------------------------------ mmcv.runner
class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()


def _passthrough_decorator(*a, **kw):
    if len(a) == 1 and callable(a[0]) and not kw:
        return a[0]
    return lambda f: f


# ------------------------------------------------------------------------------------------------ mmcv.cnn
class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def build_activation_layer(cfg):
    cfg = dict(cfg)
    t = cfg.pop('type')
    return {'ReLU': nn.ReLU, 'Swish': Swish, 'GELU': nn.GELU, 'Sigmoid': nn.Sigmoid}[t](**cfg)


def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    t = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if t in ('BN', 'BN2d'):
        name, layer = 'bn', nn.BatchNorm2d(num_features, **cfg)
    elif t == 'LN':
        name, layer = 'ln', nn.LayerNorm(num_features, **cfg)
    elif t == 'GN':
        name, layer = 'gn', nn.GroupNorm(num_channels=num_features, **cfg)
    else:
        raise KeyError(t)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return name + str(postfix), layer


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias='auto',
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True, with_spectral_norm=False,
                 padding_mode='zeros', order=('conv', 'norm', 'act')):
        super().__init__()
        assert conv_cfg is None and order == ('conv', 'norm', 'act') and not with_spectral_norm
        self.with_norm, self.with_activation = norm_cfg is not None, act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                              groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            a = dict(act_cfg)
            if a['type'] not in ('Tanh', 'PReLU', 'Sigmoid', 'HSigmoid', 'Swish', 'GELU'):
                a.setdefault('inplace', inplace)
            self.activate = build_activation_layer(a)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = getattr(self, self.norm_name)(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if getattr(module, 'weight', None) is not None:
        (nn.init.xavier_uniform_ if distribution == 'uniform' else nn.init.xavier_normal_)(module.weight, gain=gain)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def bias_init_with_prob(prior_prob):
    return float(-math.log((1 - prior_prob) / prior_prob))


class Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


# ------------------------------------------------------------------------------------------------ mmcv transformer bricks
ATTENTION = Registry('attention')
FEEDFORWARD_NETWORK = Registry('feed-forward Network')
TRANSFORMER_LAYER = Registry('transformerLayer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')
POSITIONAL_ENCODING = Registry('position encoding')
PLUGIN_LAYERS = Registry('plugin layer')


def build_dropout(cfg):
    cfg = dict(cfg)
    t = cfg.pop('type')
    assert t == 'Dropout'
    return nn.Dropout(p=cfg.pop('drop_prob', 0.5), **cfg)


@ATTENTION.register_module()
class MultiheadAttention(BaseModule):
    """mmcv/cnn/bricks/transformer.py `MultiheadAttention` (1.6.2)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=dict(type='Dropout', drop_prob=0.),
                 init_cfg=None, batch_first=False, **kwargs):
        super().__init__(init_cfg)
        dropout_layer = dict(dropout_layer) if dropout_layer else dropout_layer
        if 'dropout' in kwargs:
            attn_drop = kwargs['dropout']
            dropout_layer['drop_prob'] = kwargs.pop('dropout')
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask, key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


@FEEDFORWARD_NETWORK.register_module()
class FFN(BaseModule):
    """mmcv/cnn/bricks/transformer.py `FFN` (1.6.2)."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2
        self.embed_dims, self.feedforward_channels, self.num_fcs = embed_dims, feedforward_channels, num_fcs
        layers, c = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(c, feedforward_channels), build_activation_layer(act_cfg), nn.Dropout(ffn_drop)))
            c = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


class BaseTransformerLayer(BaseModule):
    """Imported by the reference but not instantiated on the Far3D path (Detr3DTemporalDecoderLayer derives BaseModule)."""

    def __init__(self, *a, **kw):
        raise NotImplementedError('BaseTransformerLayer is not on the Far3D path')


class TransformerLayerSequence(BaseModule):
    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        else:
            assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = nn.ModuleList()
        for i in range(num_layers):
            self.layers.append(build_from_cfg(transformerlayers[i], TRANSFORMER_LAYER))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm


def build_attention(cfg, default_args=None):
    return build_from_cfg(cfg, ATTENTION, default_args)


def build_feedforward_network(cfg, default_args=None):
    return build_from_cfg(cfg, FEEDFORWARD_NETWORK, default_args)


def build_transformer_layer_sequence(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER_SEQUENCE, default_args)


def build_positional_encoding(cfg, default_args=None):
    return build_from_cfg(cfg, POSITIONAL_ENCODING, default_args)


# ------------------------------------------------------------------------------------------------ mmcv.ops
def multi_scale_deformable_attn_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    """mmcv/ops/multi_scale_deform_attn.py: the CPU statement of `ms_deformable_im2col`.
    value (bs, sum HW, heads, dims); sampling_locations (bs, q, heads, levels, points, 2) in [0,1]; weights (bs, q, heads, levels, points)."""
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, num_heads, num_levels, num_points, _ = sampling_locations.shape
    value_list = value.split([int(H_) * int(W_) for H_, W_ in value_spatial_shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampling_value_list = []
    for level, (H_, W_) in enumerate(value_spatial_shapes):
        value_l_ = value_list[level].flatten(2).transpose(1, 2).reshape(bs * num_heads, embed_dims, int(H_), int(W_))
        sampling_grid_l_ = sampling_grids[:, :, :, level].transpose(1, 2).flatten(0, 1)
        sampling_value_list.append(F.grid_sample(value_l_, sampling_grid_l_, mode='bilinear', padding_mode='zeros',
                                                 align_corners=False))
    attention_weights = attention_weights.transpose(1, 2).reshape(bs * num_heads, 1, num_queries, num_levels * num_points)
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) * attention_weights).sum(-1) \
        .view(bs, num_heads * embed_dims, num_queries)
    return output.transpose(1, 2).contiguous()


class MultiScaleDeformableAttnFunction:
    """`.apply(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step)` with the
    CUDA op's argument layout (attention_weights (bs, q, heads, levels*points) as the reference passes it)."""

    @staticmethod
    def apply(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        bs, q, h, L, P, _ = sampling_locations.shape
        return multi_scale_deformable_attn_pytorch(value, value_spatial_shapes.tolist(), sampling_locations,
                                                   attention_weights.reshape(bs, q, h, L, P))


# ------------------------------------------------------------------------------------------------ mmdet
MODELS = Registry('models')
BACKBONES = NECKS = HEADS = DETECTORS = LOSSES = MODELS
TRANSFORMER = Registry('Transformer')
BBOX_CODERS = Registry('bbox_coder')


class _InertLoss(nn.Module):
    """Losses are built in the heads' constructors but never evaluated at inference; only `use_sigmoid` is read."""

    def __init__(self, **cfg):
        super().__init__()
        self.use_sigmoid = cfg.get('use_sigmoid', False)
        self.loss_weight = cfg.get('loss_weight', 1.0)

    def forward(self, *a, **kw):
        raise NotImplementedError('training-only')


def build_loss(cfg):
    cfg = dict(cfg)
    cfg.pop('type', None)
    return _InertLoss(**cfg)


def _training_only(*a, **kw):
    raise NotImplementedError('training-only (out of scope)')


class _Inert:
    """Assigners / samplers are constructed by the heads when a train_cfg is present but never used at inference."""

    def __init__(self, *a, **kw):
        pass


def _build_inert(*a, **kw):
    return _Inert()


def multi_apply(func, *args, **kwargs):
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def bbox_xyxy_to_cxcywh(bbox):
    x1, y1, x2, y2 = bbox.split((1, 1, 1, 1), dim=-1)
    return torch.cat([(x1 + x2) / 2, (y1 + y2) / 2, (x2 - x1), (y2 - y1)], dim=-1)


class MlvlPointGenerator:
    """mmdet/core/anchor/point_generator.py (2.28.2)."""

    def __init__(self, strides, offset=0.5):
        self.strides = [(s, s) if isinstance(s, int) else tuple(s) for s in strides]
        self.offset = offset

    @property
    def num_levels(self):
        return len(self.strides)

    @staticmethod
    def _meshgrid(x, y, row_major=True):
        yy, xx = torch.meshgrid(y, x, indexing='ij')
        return (xx.reshape(-1), yy.reshape(-1)) if row_major else (yy.reshape(-1), xx.reshape(-1))

    def grid_priors(self, featmap_sizes, dtype=torch.float32, device='cuda', with_stride=False):
        assert self.num_levels == len(featmap_sizes)
        return [self.single_level_grid_priors(featmap_sizes[i], i, dtype, device, with_stride) for i in range(self.num_levels)]

    def single_level_grid_priors(self, featmap_size, level_idx, dtype=torch.float32, device='cuda', with_stride=False):
        feat_h, feat_w = featmap_size
        stride_w, stride_h = self.strides[level_idx]
        shift_x = ((torch.arange(0, feat_w, device=device) + self.offset) * stride_w).to(dtype)
        shift_y = ((torch.arange(0, feat_h, device=device) + self.offset) * stride_h).to(dtype)
        shift_xx, shift_yy = self._meshgrid(shift_x, shift_y)
        if not with_stride:
            return torch.stack([shift_xx, shift_yy], dim=-1).to(device)
        sw = shift_xx.new_full((shift_xx.shape[0],), stride_w).to(dtype)
        sh = shift_xx.new_full((shift_yy.shape[0],), stride_h).to(dtype)
        return torch.stack([shift_xx, shift_yy, sw, sh], dim=-1).to(device)


class BaseBBoxCoder:
    def __init__(self, **kwargs):
        pass


class BaseDenseHead(BaseModule):
    def __init__(self, init_cfg=None):
        super().__init__(init_cfg)


class BBoxTestMixin:
    pass


class AnchorFreeHead(BaseDenseHead, BBoxTestMixin):
    """Only the constructor's bookkeeping: FarHead overrides `_init_layers` and calls it itself (farhead.py:222)."""

    def __init__(self, num_classes, in_channels, init_cfg=None, **kw):
        super().__init__(init_cfg)
        self.num_classes, self.in_channels = num_classes, in_channels


class NormedLinear(nn.Linear):
    def __init__(self, *a, **kw):
        raise NotImplementedError('normedlinear=False on the Far3D path')


@NECKS.register_module()
class FPN(BaseModule):
    """mmdet/models/necks/fpn.py (2.28.2)."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode='nearest'), init_cfg=None):
        super().__init__(init_cfg)
        self.in_channels, self.out_channels, self.num_ins, self.num_outs = in_channels, out_channels, len(in_channels), num_outs
        self.relu_before_extra_convs, self.no_norm_on_lateral = relu_before_extra_convs, no_norm_on_lateral
        self.upsample_cfg = dict(upsample_cfg)
        if end_level == -1 or end_level == self.num_ins - 1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level + 1
            assert end_level < self.num_ins and num_outs == end_level - start_level + 1
        self.start_level, self.end_level = start_level, end_level
        assert isinstance(add_extra_convs, (str, bool))
        if isinstance(add_extra_convs, str):
            assert add_extra_convs in ('on_input', 'on_lateral', 'on_output')
        elif add_extra_convs:
            add_extra_convs = 'on_input'
        self.add_extra_convs = add_extra_convs
        self.lateral_convs, self.fpn_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1, conv_cfg=conv_cfg,
                                                 norm_cfg=norm_cfg if not no_norm_on_lateral else None, act_cfg=act_cfg,
                                                 inplace=False))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                             act_cfg=act_cfg, inplace=False))
        extra_levels = num_outs - self.backbone_end_level + self.start_level
        if self.add_extra_convs and extra_levels >= 1:
            for i in range(extra_levels):
                c = self.in_channels[self.backbone_end_level - 1] if (i == 0 and self.add_extra_convs == 'on_input') else out_channels
                self.fpn_convs.append(ConvModule(c, out_channels, 3, stride=2, padding=1, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                                 act_cfg=act_cfg, inplace=False))

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        laterals = [l(inputs[i + self.start_level]) for i, l in enumerate(self.lateral_convs)]
        n = len(laterals)
        for i in range(n - 1, 0, -1):
            if 'scale_factor' in self.upsample_cfg:
                laterals[i - 1] = laterals[i - 1] + F.interpolate(laterals[i], **self.upsample_cfg)
            else:
                laterals[i - 1] = laterals[i - 1] + F.interpolate(laterals[i], size=laterals[i - 1].shape[2:], **self.upsample_cfg)
        outs = [self.fpn_convs[i](laterals[i]) for i in range(n)]
        if self.num_outs > len(outs):
            if not self.add_extra_convs:
                for i in range(self.num_outs - n):
                    outs.append(F.max_pool2d(outs[-1], 1, stride=2))
            else:
                if self.add_extra_convs == 'on_input':
                    src = inputs[self.backbone_end_level - 1]
                elif self.add_extra_convs == 'on_lateral':
                    src = laterals[-1]
                else:
                    src = outs[-1]
                outs.append(self.fpn_convs[n](src))
                for i in range(n + 1, self.num_outs):
                    outs.append(self.fpn_convs[i](F.relu(outs[-1]) if self.relu_before_extra_convs else outs[-1]))
        return tuple(outs)


# ------------------------------------------------------------------------------------------------ mmdet3d
class MVXTwoStageDetector(BaseModule):
    """mmdet3d/models/detectors/mvx_two_stage.py: only the construction of the image-branch sub-modules and heads."""

    def __init__(self, pts_voxel_layer=None, pts_voxel_encoder=None, pts_middle_encoder=None, pts_fusion_layer=None,
                 img_backbone=None, pts_backbone=None, img_neck=None, pts_neck=None, pts_bbox_head=None, img_roi_head=None,
                 img_rpn_head=None, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__(init_cfg)
        assert not any((pts_voxel_layer, pts_voxel_encoder, pts_middle_encoder, pts_fusion_layer, pts_backbone, pts_neck,
                        img_rpn_head))
        if pts_bbox_head:
            pts_bbox_head = dict(pts_bbox_head)
            pts_bbox_head.update(train_cfg=train_cfg.pts if train_cfg else None)
            pts_bbox_head.update(test_cfg=test_cfg.pts if test_cfg else None)
            self.pts_bbox_head = build_from_cfg(pts_bbox_head, HEADS)
        if img_backbone:
            self.img_backbone = build_from_cfg(img_backbone, BACKBONES)
        if img_neck is not None:
            self.img_neck = build_from_cfg(img_neck, NECKS)
        if img_roi_head is not None:
            self.img_roi_head = build_from_cfg(img_roi_head, HEADS)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

    @property
    def with_img_neck(self):
        return hasattr(self, 'img_neck') and self.img_neck is not None

    @property
    def with_img_roi_head(self):
        return hasattr(self, 'img_roi_head') and self.img_roi_head is not None


def bbox3d2result(bboxes, scores, labels, attrs=None):
    return dict(boxes_3d=bboxes.to('cpu'), scores_3d=scores.cpu(), labels_3d=labels.cpu())


class Boxes3D:
    """Minimal `box_type_3d` for img_metas (LiDARInstance3DBoxes only wraps the (K, box_dim) tensor on this path)."""

    def __init__(self, tensor, box_dim=7):
        self.tensor, self.box_dim = tensor, box_dim

    def to(self, device):
        return Boxes3D(self.tensor.to(device), self.box_dim)


# ------------------------------------------------------------------------------------------------ mmcv.image (test pipeline)
def imnormalize(img, mean, std, to_rgb=True):
    """mmcv-full 1.6.2 mmcv/image/photometric.py imnormalize / imnormalize_: float32 copy, optional BGR->RGB, then
    cv2.subtract(img, float64(mean)) and cv2.multiply(img, 1 / float64(std)), which OpenCV evaluates in float32 on a float32
    image (restated with numpy; cv2 is not installed here)."""
    import numpy as np
    img = img.copy().astype(np.float32)
    assert img.dtype != np.uint8
    mean = np.float64(np.asarray(mean).reshape(1, -1))
    stdinv = 1 / np.float64(np.asarray(std).reshape(1, -1))
    if to_rgb:
        img = np.ascontiguousarray(img[..., ::-1])
    img = (img - mean.astype(np.float32)).astype(np.float32)
    return (img * stdinv.astype(np.float32)).astype(np.float32)


def impad(img, *, shape=None, padding=None, pad_val=0, padding_mode='constant'):
    """mmcv/image/geometric.py impad with `shape`: pad bottom / right with pad_val (cv2.copyMakeBorder, BORDER_CONSTANT)."""
    import numpy as np
    assert shape is not None and padding is None and padding_mode == 'constant'
    h, w = img.shape[:2]
    out = np.full((shape[0], shape[1]) + img.shape[2:], pad_val, dtype=img.dtype)
    out[:h, :w] = img
    return out


def impad_to_multiple(img, divisor, pad_val=0):
    import numpy as np
    h = int(np.ceil(img.shape[0] / divisor)) * divisor
    w = int(np.ceil(img.shape[1] / divisor)) * divisor
    return impad(img, shape=(h, w), pad_val=pad_val)


def load_reference_pipelines(root='/root/reference'):
    """The reference's own image transforms (datasets/pipelines/transform_3d.py, custom_pipeline.py), loaded by file with the
    third-party / site-specific imports they make at module level stubbed (mmdet PIPELINES registry, mmdet3d points, refile)."""
    import importlib.util
    import os
    install()
    PIPELINES = Registry('pipeline')
    _mod('mmcv', imnormalize=imnormalize, impad=impad, impad_to_multiple=impad_to_multiple)
    _mod('mmdet.datasets')
    _mod('mmdet.datasets.builder', PIPELINES=PIPELINES)
    _mod('mmdet3d.datasets')
    _mod('mmdet3d.datasets.builder', PIPELINES=PIPELINES)
    _mod('mmdet3d.core.points', BasePoints=object, get_points_type=lambda *a, **k: None)
    _mod('refile')
    plug = os.path.join(root, 'projects', 'mmdet3d_plugin', 'datasets', 'pipelines')
    out = {}
    for rel in ('transform_3d.py', 'custom_pipeline.py'):
        name = 'far3d_ref_pipelines.' + rel[:-3]
        spec = importlib.util.spec_from_file_location(name, os.path.join(plug, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        out[rel] = m
    return out


def load_reference_av2_export(root='/root/reference'):
    """The reference's own `Argoverse2Dataset.format_results` / `box_to_av2` (datasets/argoverse2_dataset.py) and `yaw_to_quat`
    (datasets/av2_utils.py), loaded by file.  Their module-level imports of packages that are absent offline (av2, kornia, the
    mmdet / mmdet3d dataset bases, the evaluation helper) are satisfied by inert stand-ins: none of them is executed by the
    export path except `LiDARInstance3DBoxes.gravity_center`, restated below as mmdet3d defines it."""
    import importlib.util
    import os
    import types

    class _AnyMeta(type):                                   # a class whose every attribute exists (enum members, constants)
        def __getattr__(cls, k):
            if k.startswith('__'):
                raise AttributeError(k)
            return 0

    class _Anything(types.ModuleType):                      # a module whose every attribute is such a class
        def __getattr__(self, k):
            if k.startswith('__'):
                raise AttributeError(k)
            return _AnyMeta(k, (), {})

    install()
    for name in ('av2', 'av2.evaluation', 'av2.evaluation.detection', 'av2.evaluation.detection.constants', 'av2.geometry',
                 'av2.geometry.geometry', 'av2.geometry.iou', 'av2.geometry.se3', 'av2.map', 'av2.map.map_api', 'av2.structures',
                 'av2.structures.cuboid', 'av2.utils', 'av2.utils.typing', 'av2.utils.io', 'kornia', 'kornia.geometry',
                 'kornia.geometry.conversions'):
        m = _Anything(name); m.__path__ = []
        sys.modules[name] = m
        if '.' in name:
            setattr(sys.modules[name.rsplit('.', 1)[0]], name.rsplit('.', 1)[1], m)
    ns = {}
    with open(os.path.join(root, 'projects', 'configs', 'far3d.py')) as f:
        exec(compile(f.read(), 'far3d.py', 'exec'), ns)
    import enum
    sys.modules['av2.evaluation.detection.constants'].CompetitionCategories = enum.Enum(
        'CompetitionCategories', {c: c for c in ns['class_names']})        # the config's own class list, in its order
    for k in ('MAX_NORMALIZED_ASE', 'MAX_SCALE_ERROR', 'MAX_YAW_RAD_ERROR', 'MIN_AP', 'MIN_CDS'):
        setattr(sys.modules['av2.evaluation.detection.constants'], k, 0.0)

    class LiDARInstance3DBoxes:
        """mmdet3d/core/bbox/structures: (K, 7) tensor, bottom-centre origin; gravity_center adds h / 2 on z"""

        def __init__(self, tensor, box_dim=7):
            self.tensor = tensor

        @property
        def gravity_center(self):
            bc = self.tensor[:, :3]
            gc = torch.zeros_like(bc)
            gc[:, :2] = bc[:, :2]
            gc[:, 2] = bc[:, 2] + self.tensor[:, 5] * 0.5
            return gc

    _mod('mmdet3d.core.bbox', LiDARInstance3DBoxes=LiDARInstance3DBoxes)
    _mod('mmdet.datasets', DATASETS=Registry('dataset'))
    _mod('mmdet3d.datasets')
    _mod('mmdet3d.datasets.custom_3d', Custom3DDataset=object)
    _mod('mmcv', track_iter_progress=lambda it: it, mkdir_or_exist=lambda p: None)
    pkg = _mod('far3d_ref_datasets')
    base = os.path.join(root, 'projects', 'mmdet3d_plugin', 'datasets')
    _mod('far3d_ref_datasets.av2_eval_util', evaluate=None)
    out = {}
    for rel in ('av2_utils.py', 'argoverse2_dataset.py'):
        name = 'far3d_ref_datasets.' + rel[:-3]
        spec = importlib.util.spec_from_file_location(name, os.path.join(base, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        setattr(pkg, rel[:-3], m)
        out[rel] = m
    out['LiDARInstance3DBoxes'] = LiDARInstance3DBoxes
    out['class_names'] = ns['class_names']
    return out


# ------------------------------------------------------------------------------------------------ installation
def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []          # behaves as a package for sub-module imports
        sys.modules[name] = m
        if '.' in name:
            parent, leaf = name.rsplit('.', 1)
            setattr(_mod(parent), leaf, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install():
    """Registers the stand-in packages in sys.modules (refuses to shadow a real installation)."""
    for top in ('mmcv', 'mmdet', 'mmdet3d'):
        if top in sys.modules and not getattr(sys.modules[top], '_far3d_shim', False):
            raise RuntimeError(f'{top} is really installed; use it instead of the shims')
    for top in ('mmcv', 'mmdet', 'mmdet3d'):
        _mod(top, _far3d_shim=True)
    _mod('mmcv.utils', Registry=Registry, build_from_cfg=build_from_cfg, ConfigDict=ConfigDict,
         deprecated_api_warning=deprecated_api_warning)
    _mod('mmcv.runner', BaseModule=BaseModule, force_fp32=_passthrough_decorator, auto_fp16=_passthrough_decorator)
    _mod('mmcv.runner.base_module', BaseModule=BaseModule)
    _mod('mmcv.cnn', ConvModule=ConvModule, DepthwiseSeparableConvModule=_training_only, Linear=nn.Linear, Scale=Scale,
         bias_init_with_prob=bias_init_with_prob, xavier_init=xavier_init, constant_init=constant_init,
         build_norm_layer=build_norm_layer)
    _mod('mmcv.cnn.bricks')
    _mod('mmcv.cnn.bricks.registry', ATTENTION=ATTENTION, TRANSFORMER_LAYER=TRANSFORMER_LAYER,
         TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE, PLUGIN_LAYERS=PLUGIN_LAYERS,
         POSITIONAL_ENCODING=POSITIONAL_ENCODING, FEEDFORWARD_NETWORK=FEEDFORWARD_NETWORK)
    _mod('mmcv.cnn.bricks.transformer', BaseTransformerLayer=BaseTransformerLayer, TransformerLayerSequence=TransformerLayerSequence,
         build_transformer_layer_sequence=build_transformer_layer_sequence, build_attention=build_attention,
         build_feedforward_network=build_feedforward_network, build_positional_encoding=build_positional_encoding,
         POSITIONAL_ENCODING=POSITIONAL_ENCODING, FFN=FFN, MultiheadAttention=MultiheadAttention)
    _mod('mmcv.ops')
    _mod('mmcv.ops.multi_scale_deform_attn', MultiScaleDeformableAttnFunction=MultiScaleDeformableAttnFunction,
         multi_scale_deformable_attn_pytorch=multi_scale_deformable_attn_pytorch)
    _mod('mmcv.ops.nms', batched_nms=_training_only)
    _mod('mmdet.models', HEADS=HEADS, DETECTORS=DETECTORS, build_loss=build_loss)
    _mod('mmdet.models.builder', BACKBONES=BACKBONES, NECKS=NECKS, HEADS=HEADS, DETECTORS=DETECTORS, build_loss=build_loss)
    _mod('mmdet.models.utils', build_transformer=lambda cfg, default_args=None: build_from_cfg(cfg, TRANSFORMER, default_args),
         NormedLinear=NormedLinear)
    _mod('mmdet.models.utils.builder', TRANSFORMER=TRANSFORMER)
    _mod('mmdet.models.utils.transformer', inverse_sigmoid=inverse_sigmoid)
    _mod('mmdet.models.dense_heads')
    _mod('mmdet.models.dense_heads.anchor_free_head', AnchorFreeHead=AnchorFreeHead)
    _mod('mmdet.models.dense_heads.base_dense_head', BaseDenseHead=BaseDenseHead)
    _mod('mmdet.models.dense_heads.dense_test_mixins', BBoxTestMixin=BBoxTestMixin)
    _mod('mmdet.core', build_assigner=_build_inert, build_sampler=_build_inert, multi_apply=multi_apply,
         reduce_mean=_training_only, MlvlPointGenerator=MlvlPointGenerator, bbox_xyxy_to_cxcywh=bbox_xyxy_to_cxcywh)
    _mod('mmdet.core.bbox', BaseBBoxCoder=BaseBBoxCoder)
    _mod('mmdet.core.bbox.builder', BBOX_CODERS=BBOX_CODERS)
    _mod('mmdet3d.core', bbox3d2result=bbox3d2result)
    _mod('mmdet3d.core.bbox')
    _mod('mmdet3d.core.bbox.coders', build_bbox_coder=lambda cfg, **kw: build_from_cfg(cfg, BBOX_CODERS))
    _mod('mmdet3d.models')
    _mod('mmdet3d.models.detectors')
    _mod('mmdet3d.models.detectors.mvx_two_stage', MVXTwoStageDetector=MVXTwoStageDetector)


REF_FILES = [
    # the reference's own modules on the inference path, in dependency order; loaded by file so that no package
    # __init__ (datasets, training hooks, ...) runs
    'core/bbox/util.py',
    'core/bbox/coders/nms_free_coder.py',
    'models/utils/positional_encoding.py',
    'models/utils/misc.py',
    'models/utils/grid_mask.py',
    'models/utils/detr3d_transformer.py',
    'models/backbones/vovnet.py',
    'models/depth_predictor/depth_predictor.py',
    'models/dense_heads/yolox_head.py',
    'models/dense_heads/farhead.py',
    'models/detectors/far3d.py',
]


def load_reference(root='/root/reference'):
    """Imports the reference's own files (unmodified, from where they lie) under their real module names."""
    import importlib.util
    import os
    install()
    plug = os.path.join(root, 'projects', 'mmdet3d_plugin')
    if not os.path.isdir(plug):
        raise FileNotFoundError(plug)
    _mod('projects')
    _mod('projects.mmdet3d_plugin')
    # yolox_head.py:21-22 imports the depth_predictor package relatively; DDNLoss is a training loss (inert stand-in)
    out = {}
    for rel in REF_FILES:
        name = 'projects.mmdet3d_plugin.' + rel[:-3].replace('/', '.')
        parent = name.rsplit('.', 1)[0]
        _mod(parent)
        if rel == 'models/depth_predictor/depth_predictor.py':
            pkg = _mod('projects.mmdet3d_plugin.models.depth_predictor')
        spec = importlib.util.spec_from_file_location(name, os.path.join(plug, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        setattr(sys.modules[parent], name.rsplit('.', 1)[1], m)
        if rel == 'models/depth_predictor/depth_predictor.py':
            pkg.DepthPredictor = m.DepthPredictor
            _mod('projects.mmdet3d_plugin.models.depth_predictor.ddn_loss', DDNLoss=lambda *a, **k: None)
        out[rel] = m
    return out


def reference_model_cfg(root='/root/reference'):
    """The `model` dict of the reference's projects/configs/far3d.py, evaluated from the file itself."""
    import os
    ns = {}
    with open(os.path.join(root, 'projects', 'configs', 'far3d.py')) as f:
        exec(compile(f.read(), 'far3d.py', 'exec'), ns)
    return ns['model']
