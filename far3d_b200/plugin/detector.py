"""`Far3D` detector (test path) wired from the reference's config dict.

Reference: projects/mmdet3d_plugin/models/detectors/far3d.py - extract_img_feat :64-99, forward :166-180,
forward_test :232-242, simple_test_pts :244-266, simple_test :268-277.  Training entry points raise.
mmdet3d's `MVXTwoStageDetector` base (third-party) is replaced by the few attributes the test path reads."""
import torch
import torch.nn as nn

from .. import _lib, ops
from ..compat import BACKBONES, DETECTORS, HEADS, NECKS, build_from_cfg


def _drop_image_graphs(module, incompatible_keys):
    module.__dict__.pop('_img_graphs', None)         # weights changed: the captured launches hold stale packed copies


@DETECTORS.register_module()
class Far3D(nn.Module):
    def __init__(self, use_grid_mask=False, pts_voxel_layer=None, pts_voxel_encoder=None, pts_middle_encoder=None,
                 pts_fusion_layer=None, img_backbone=None, pts_backbone=None, img_neck=None, pts_neck=None,
                 pts_bbox_head=None, img_roi_head=None, img_rpn_head=None, train_cfg=None, test_cfg=None,
                 depth_branch=None, stride=[16], position_level=[0], aux_2d_only=True, single_test=False,
                 pretrained=None):
        super().__init__()
        assert depth_branch is None, 'depth_branch is not used by far3d.py'
        self.img_backbone = build_from_cfg(img_backbone, BACKBONES) if img_backbone else None
        self.img_neck = build_from_cfg(img_neck, NECKS) if img_neck is not None else None
        if pts_bbox_head is not None:
            cfg = dict(pts_bbox_head)
            cfg.update(train_cfg=(train_cfg or {}).get('pts') if train_cfg else None,
                       test_cfg=(test_cfg or {}).get('pts') if test_cfg else None)
            self.pts_bbox_head = build_from_cfg(cfg, HEADS)
        else:
            self.pts_bbox_head = None
        self.img_roi_head = build_from_cfg(img_roi_head, HEADS) if img_roi_head is not None else None
        self.use_grid_mask = use_grid_mask          # GridMask is a no-op outside training (grid_mask.py:83-85)
        self.prev_scene_token = None
        self.single_test, self.stride, self.position_level = single_test, list(stride), list(position_level)
        self.aux_2d_only = aux_2d_only
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.section_events = None      # bench hook: list that receives (name, cuda event) marks per frame
        # mmdet3d's bbox3d2result (what the reference's simple_test_pts returns, far3d.py:262-265) moves boxes / scores / labels
        # to the host; tools/test.py and Argoverse2Dataset.format_results call .numpy() on them.  far3d_b200 extension: True keeps
        # them on the device (Far3DPipeline does its own D2H).
        self.results_on_device = False
        self.register_load_state_dict_post_hook(_drop_image_graphs)

    def _mark(self, name):
        if self.section_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.section_events.append((name, e))

    @property
    def with_img_neck(self):
        return self.img_neck is not None

    @property
    def with_img_roi_head(self):
        return self.img_roi_head is not None

    def init_weights(self):
        for m in (self.pts_bbox_head, self.img_roi_head):
            if m is not None and hasattr(m, 'init_weights'):
                m.init_weights()

    def set_precision(self, precision):
        """far3d_b200 extension: 'fp16x3' (fp32-grade, three tensor passes), 'fp16mx' (fp16 + e4m3 correction stream, two
        passes), 'fp16' (fastest) or 'fp32' (SIMT anchor)."""
        self.__dict__.pop('_img_graphs', None)
        for m in self.modules():
            if m is not self and hasattr(m, 'set_precision'):
                m.set_precision(precision)

    def extract_img_feat(self, img, return_depth=False):
        B = img.size(0)
        if img.dim() == 6:
            img = img.flatten(1, 2)
        if img.dim() == 5:
            B, N, C, H, W = img.size()
            img = img.reshape(B * N, C, H, W)
        feats = self.img_backbone(img)
        if isinstance(feats, dict):
            feats = list(feats.values())
        if self.with_img_neck:
            feats = self.img_neck(feats)
        out = []
        for i in self.position_level:
            BN, C, H, W = feats[i].size()
            v = feats[i].view(B, BN // B, C, H, W)
            buf = getattr(feats[i], '_far3d_buf', None)
            if buf is not None:
                v._far3d_buf = buf          # NHWC planes ride along for the heads (zero-copy)
            out.append(v)
        return (out, None) if return_depth else out

    def extract_feat(self, img, return_depth=False):
        return self.extract_img_feat(img, return_depth)

    # ------------------------------------------------------------------ static image branch as one CUDA graph
    use_cuda_graph = True      # backbone + FPN + 2D-head convolutions: ~260 launches with fixed shapes

    def _image_branch_eager(self, img):
        feats = self.extract_img_feat(img)
        outs_roi = self.img_roi_head(None, img_feats=feats) if self.with_img_roi_head else None
        return feats, outs_roi

    def _image_branch_graph(self, img, slot=0):
        """replay (capture on first use) the image branch; `slot` selects one of several graph instances with their own input
        and output buffers, so a pipelined caller can run frame i+1's image branch while frame i's outputs are still read."""
        shape_key = (tuple(img.shape), str(img.device), ops.LINEAR_MODE,
                     tuple(getattr(m, 'precision', None) for m in (self.img_backbone, self.img_neck, self.img_roi_head)))
        key = shape_key + (slot,)
        cache = self.__dict__.setdefault('_img_graphs', {})
        ent = cache.get(key)
        if ent is None:
            for k in [k for k in cache if k[:-1] != shape_key]:
                del cache[k]                            # one input shape at a time: the activation plans are per shape too
            static = img.detach().clone().contiguous()
            if self.with_img_neck:
                self.img_neck.plan_slot = slot          # this instance's own FPN output maps (baked into the capture below)
            self._image_branch_eager(static)            # warm-up: packs weights, builds buffer plans
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g):
                outs = self._image_branch_eager(static)
            ent = cache[key] = (g, static, outs, _lib.launch_count() - n0)
        g, static, outs, nlaunch = ent
        static.copy_(img)
        g.replay()
        _lib.load().far3d_add_launches(nlaunch)
        return outs

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop('_img_graphs', None)
        return super()._apply(fn, *args, **kwargs)

    def forward(self, return_loss=True, **data):
        if return_loss:
            raise NotImplementedError('far3d_b200 implements the inference path (return_loss=False) only')
        return self.forward_test(**data)

    def forward_train(self, *a, **k):
        raise NotImplementedError('training is out of scope for far3d_b200 (SURVEY.md section 2)')

    def forward_test(self, img_metas, rescale=False, **data):
        if not isinstance(img_metas, list):
            raise TypeError('img_metas must be a list, but got {}'.format(type(img_metas)))
        for key in data:
            if key not in ['img', 'gt_bboxes_3d', 'gt_bboxes', 'centers2d']:
                data[key] = data[key][0][0].unsqueeze(0)
            else:
                data[key] = data[key][0]
        return self.simple_test(img_metas[0], **data)

    def simple_test_pts(self, img_metas, **data):
        outs_roi = data.pop('_outs_roi_dense', None)
        if self.with_img_roi_head:
            if outs_roi is None:
                outs_roi = self.img_roi_head(None, **data)
                self._mark('roi_head_convs')
            outs_roi = dict(outs_roi)
            sel = None
            if (getattr(self.pts_bbox_head, 'proposal_kernels', False) and data['lidar2img'].shape[0] == 1
                    and outs_roi.get('_depth_logit_nhwc') is not None):
                sel = self.img_roi_head.select_device(outs_roi)       # sync-free peak pick + compaction
            if sel is not None:
                outs_roi['_sel'] = sel
            else:
                outs_roi.update(self.img_roi_head.get_bboxes(outs_roi))
            self._mark('roi_proposals')
        inj = data.pop('inject_roi', None)
        if inj is not None:
            outs_roi = dict(outs_roi or {}, **inj)
        if img_metas[0]['scene_token'] != self.prev_scene_token:
            self.prev_scene_token = img_metas[0]['scene_token']
            data['prev_exists'] = data['img'].new_zeros(1)
            self.pts_bbox_head.reset_memory()
        else:
            data['prev_exists'] = data['img'].new_ones(1)
        outs = self.pts_bbox_head(img_metas, outs_roi, **data)
        self._mark('pts_head')
        self.last_outs = outs
        bbox_list = self.pts_bbox_head.get_bboxes(outs, img_metas)
        self._mark('decode')
        if not self.results_on_device:
            bbox_list = [[(b.to('cpu') if hasattr(b, 'to') else b), s.cpu(), l.cpu()] for b, s, l in bbox_list]
        results = [dict(boxes_3d=b, scores_3d=s, labels_3d=l) for b, s, l in bbox_list]
        return results, (outs_roi or {}).get('bbox_list')

    @torch.no_grad()
    def image_branch(self, img, slot=0):
        """backbone + FPN + 2D-head convolutions of one frame on the current stream -> (img_feats, dense 2D-head outputs)."""
        if self.use_cuda_graph and ops.PROFILE is None and img.is_cuda:
            out = self._image_branch_graph(img, slot)
            self._mark('image_branch_graph')
            return out
        feats = self.extract_img_feat(img)
        self._mark('backbone_fpn')
        return feats, None

    @torch.no_grad()
    def simple_test(self, img_metas, **data):
        self._mark('start')
        if '_img_feats' in data:                        # pipelined caller (Far3DPipeline.submit / collect) ran the image branch already
            data['img_feats'], data['_outs_roi_dense'] = data.pop('_img_feats')
        else:
            data['img_feats'], data['_outs_roi_dense'] = self.image_branch(data['img'])
        bbox_list = [dict() for _ in range(len(img_metas))]
        bbox_pts, _ = self.simple_test_pts(img_metas, **data)
        for r, p in zip(bbox_list, bbox_pts):
            r['pts_bbox'] = p
        return bbox_list
