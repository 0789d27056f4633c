"""2D proposal head + depth classifier on the sm_100a conv kernels (registry name `YOLOXHeadCustom`).

Reference: projects/mmdet3d_plugin/models/dense_heads/yolox_head.py (layers :164-231, forward :260-341, get_bboxes
:355-489, box decode :491-501) and models/depth_predictor/depth_predictor.py:38-86.  Test path only.

Convolutions (towers with folded BN + Swish, 1x1 predictors, depth head with GroupNorm + ReLU) run on the tensor-core
conv kernel; the data-dependent proposal selection (3x3 local-max, score threshold, boolean gather) is torch device glue
(SURVEY.md section 8 row f1 marks its sync-free rewrite as the next step)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..compat import HEADS
from .backbone import PRECISIONS, Buf, PackedConv, PackedMixin, _as_buf, fold_bn, run_conv


class _ConvBNAct(nn.Module):
    """mmcv ConvModule(conv bias=False, BN(eps=1e-3, momentum=0.03), Swish): children `conv`, `bn`."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, 1, 1, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=0.001, momentum=0.03)


class DepthPredictor(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        d = 256
        mk = lambda a, b: nn.Sequential(nn.Conv2d(a, b, kernel_size=(3, 3), padding=1), nn.GroupNorm(32, b), nn.ReLU())
        assert not model_cfg.get('multi_level_fusion', False)
        n = model_cfg.get('conv_layer_num', 2)
        self.depth_head = nn.Sequential(mk(int(model_cfg['hidden_dim']), d), *[mk(d, d) for _ in range(n - 1)])
        self.depth_classifier = nn.Conv2d(d, int(model_cfg['num_depth_bins']) + 1, kernel_size=(1, 1))


@HEADS.register_module()
class YOLOXHeadCustom(PackedMixin, nn.Module):
    def __init__(self, num_classes, in_channels, feat_channels=256, stacked_convs=2, strides=[8, 16, 32],
                 use_depthwise=False, dcn_on_last_conv=False, conv_bias='auto', conv_cfg=None,
                 norm_cfg=dict(type='BN', momentum=0.03, eps=0.001), act_cfg=dict(type='Swish'), train_cfg=None,
                 test_cfg=None, init_cfg=None, pred_with_depth=False, depthnet_config={}, reg_depth_level='p4',
                 pred_depth_var=False, sample_with_score=True, threshold_score=0.05, topk_proposal=None,
                 return_context_feat=False, embedding_cam=False, precision='fp16x3', **training_only):
        super().__init__()
        assert not use_depthwise and not dcn_on_last_conv and not pred_depth_var and not embedding_cam
        assert act_cfg.get('type') == 'Swish' and sample_with_score
        self.num_classes = self.cls_out_channels = num_classes
        self.in_channels, self.feat_channels, self.strides = in_channels, feat_channels, list(strides)
        self.threshold_score, self.sample_with_score = threshold_score, sample_with_score
        self.pred_with_depth, self.reg_depth_level = pred_with_depth, reg_depth_level
        self.depthnet_config = dict(depthnet_config)
        self.return_context_feat = return_context_feat
        self.test_cfg, self.train_cfg = test_cfg, train_cfg
        mk = lambda: nn.Sequential(*[_ConvBNAct(in_channels if i == 0 else feat_channels, feat_channels)
                                     for i in range(stacked_convs)])
        self.multi_level_cls_convs = nn.ModuleList(mk() for _ in strides)
        self.multi_level_reg_convs = nn.ModuleList(mk() for _ in strides)
        self.multi_level_conv_cls = nn.ModuleList(nn.Conv2d(feat_channels, num_classes, 1) for _ in strides)
        self.multi_level_conv_reg = nn.ModuleList(nn.Conv2d(feat_channels, 4, 1) for _ in strides)
        self.multi_level_conv_obj = nn.ModuleList(nn.Conv2d(feat_channels, 1, 1) for _ in strides)
        self.multi_level_conv_centers2d = nn.ModuleList(nn.Conv2d(feat_channels, 2, 1) for _ in strides)
        if pred_with_depth:
            assert self.depthnet_config.get('type', 0) == 0 and not self.depthnet_config.get('multi_level_pred', False)
            self.depthnet = DepthPredictor(self.depthnet_config)
        self.precision = precision
        self._packed = None
        self._plan = None

    def init_weights(self):                       # yolox_head.py:226-236 (kaiming-uniform convs + prior-prob biases)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=math.sqrt(5), mode='fan_in', nonlinearity='leaky_relu')
        b = float(-math.log((1 - 0.01) / 0.01))
        for c, o in zip(self.multi_level_conv_cls, self.multi_level_conv_obj):
            c.bias.data.fill_(b); o.bias.data.fill_(b)

    def set_precision(self, precision):
        assert precision in PRECISIONS
        if precision != self.precision:
            self.precision, self._packed = precision, None

    @torch.no_grad()
    def _pack(self):
        pr = self.precision
        pk = dict(cls=[], reg=[], pred_cls=[], pred_reg=[])
        for i in range(len(self.strides)):
            pk['cls'].append([PackedConv(*fold_bn(m.conv.weight, m.bn), pr) for m in self.multi_level_cls_convs[i]])
            pk['reg'].append([PackedConv(*fold_bn(m.conv.weight, m.bn), pr) for m in self.multi_level_reg_convs[i]])
            c = self.multi_level_conv_cls[i]
            pk['pred_cls'].append(PackedConv(c.weight, c.bias, pr))
            r, o, t = self.multi_level_conv_reg[i], self.multi_level_conv_obj[i], self.multi_level_conv_centers2d[i]
            pk['pred_reg'].append(PackedConv(torch.cat([r.weight, o.weight, t.weight], 0),
                                             torch.cat([r.bias, o.bias, t.bias], 0), pr))     # 4 + 1 + 2 outputs
        if self.pred_with_depth:
            pk['depth'] = [PackedConv(blk[0].weight, blk[0].bias, pr) for blk in self.depthnet.depth_head]
            pk['depth_cls'] = PackedConv(self.depthnet.depth_classifier.weight, self.depthnet.depth_classifier.bias, pr)
        self._packed = pk

    @torch.no_grad()
    def forward(self, locations=None, **data):
        if self.training:
            raise RuntimeError('far3d_b200 YOLOXHeadCustom implements the test path only; call .eval()')
        self._check_packed()
        if self._packed is None:
            self._pack()
        pr, pk = self.precision, self._packed
        feats = data['img_feats']
        cls_scores, bbox_preds, objs, ctrs = [], [], [], []
        srcs = []
        nhwc = ([], [])                                  # the predictors' own NHWC buffers (cls | box + obj + centre), per level
        for i, f in enumerate(feats):
            x = f.flatten(0, 1) if f.dim() == 5 else f
            src = _as_buf(_carry(f, x), pr)
            srcs.append(src)
            N, H, W, dev = src.N, src.H, src.W, x.device
            outs = []
            for tower, pred, cout in ((pk['cls'][i], pk['pred_cls'][i], self.num_classes), (pk['reg'][i], pk['pred_reg'][i], 7)):
                cur = src
                for pc in tower:
                    nxt = Buf(N, H, W, pc.Cout, dev, pr)
                    run_conv(pc, pr, cur, 0, dst_f32=nxt if pr == 'fp32' else None, dst_b=nxt if pr != 'fp32' else None,
                             relu=2)
                    cur = nxt
                o = Buf(N, H, W, cout if cout % 4 == 0 else (cout + 3) // 4 * 4, dev, 'fp32')
                run_conv(pred, pr, cur, 0, dst_f32=o, relu=False)
                outs.append(o.f32.permute(0, 3, 1, 2))
                nhwc[0 if cout == self.num_classes else 1].append(o.f32)
            cls_scores.append(outs[0][:, :self.num_classes])
            bbox_preds.append(outs[1][:, 0:4]); objs.append(outs[1][:, 4:5]); ctrs.append(outs[1][:, 5:7])
        out = dict(enc_cls_scores=cls_scores, enc_bbox_preds=bbox_preds, pred_centers2d_offset=ctrs, objectnesses=objs,
                   topk_indexes=None, _cls_nhwc=nhwc[0], _reg_nhwc=nhwc[1])
        if self.pred_with_depth:
            ridx = ['p3', 'p4', 'p5'].index(self.reg_depth_level)          # yolox_head.py:300-301
            cur = srcs[ridx]
            N, H, W, dev = cur.N, cur.H, cur.W, feats[ridx].device
            for pc, blk in zip(pk['depth'], self.depthnet.depth_head):
                t = Buf(N, H, W, pc.Cout, dev, 'fp32')
                run_conv(pc, pr, cur, 0, dst_f32=t, relu=False)
                gn = blk[1]
                nxt = Buf(N, H, W, pc.Cout, dev, pr)
                ops.groupnorm_nhwc(t.f32, gn.weight, gn.bias, N, H * W, pc.Cout, gn.num_groups, gn.eps, True,
                                   y_f32=nxt.f32, y_hi=nxt.hi, y_lo=nxt.lo, lo_fmt=nxt.fmt)
                cur = nxt
            nb = pk['depth_cls'].Cout
            lg = Buf(N, H, W, (nb + 3) // 4 * 4, dev, 'fp32')
            run_conv(pk['depth_cls'], pr, cur, 0, dst_f32=lg, relu=False)
            logit = lg.f32.permute(0, 3, 1, 2)[:, :nb]
            out.update(depth_logit=logit, pred_depth=logit.softmax(dim=1), _depth_logit_nhwc=lg.f32, _depth_bins=nb)
        return out

    def _priors(self, sizes, device):
        res = []
        for (h, w), s in zip(sizes, self.strides):       # mmdet MlvlPointGenerator(strides, offset=0), with_stride=True
            xs = torch.arange(0, w, device=device, dtype=torch.float32) * s
            ys = torch.arange(0, h, device=device, dtype=torch.float32) * s
            xx = xs.repeat(h); yy = ys.view(-1, 1).repeat(1, w).view(-1)
            res.append(torch.stack([xx, yy, xx.new_full(xx.shape, s), xx.new_full(xx.shape, s)], dim=-1))
        return res

    select_cap = 1024          # per-camera capacity of the device-side proposal slots (cfg-5 peaks at ~150 per camera)

    @torch.no_grad()
    def select_device(self, preds_dicts):
        """get_bboxes on the device, sync-free: per-camera slots of peaks in the reference's order (far3d_roi_select).
        None when the dense maps did not come from this module's own NHWC buffers (e.g. gathered from other ranks)."""
        cls, reg = preds_dicts.get('_cls_nhwc'), preds_dicts.get('_reg_nhwc')
        if not cls or not reg or len(cls) != len(self.strides):
            return None
        return ops.roi_select(cls, reg, self.strides, self.num_classes, self.threshold_score, self.select_cap)

    @torch.no_grad()
    def get_bboxes(self, preds_dicts, img_metas=None, cfg=None, rescale=False, with_nms=True, threshold_score=0.1, **data):
        cls, box, obj = preds_dicts['enc_cls_scores'], preds_dicts['enc_bbox_preds'], preds_dicts['objectnesses']
        n = cls[0].shape[0]
        pri = torch.cat(self._priors([c.shape[2:] for c in cls], cls[0].device))
        sw = []
        for i in range(len(obj)):
            w = obj[i].sigmoid() * cls[i].amax(dim=1, keepdim=True).sigmoid()       # == topk(1).values (yolox_head.py:455)
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
            b = boxes[i][valid[i, :, 0]]                    # row gather == masked_select with the mask repeated over 4 columns
            res.append(torch.cat([(b[:, :2] + b[:, 2:]) / 2, b[:, 2:] - b[:, :2]], dim=-1))
        return dict(bbox_list=res, bbox2d_scores=score[valid].reshape(-1, 1), valid_indices=valid)


def _carry(orig, flat):
    b = getattr(orig, '_far3d_buf', None)
    if b is not None:
        flat._far3d_buf = b
    return flat
