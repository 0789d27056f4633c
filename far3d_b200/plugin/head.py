"""FarHead (inference path), MLN and NMSFreeCoder on the sm_100a kernels.

Reference: projects/mmdet3d_plugin/models/dense_heads/farhead.py (forward :533-693, memory bank :446-508, temporal
alignment :284-313, 2D->3D query lifting :710-827, get_bboxes :1224-1245), models/utils/misc.py (MLN :153-190),
core/bbox/coders/nms_free_coder.py:39-112, core/bbox/util.py:25-52.  Same registry names, ctor arguments and
state_dict keys; training-only arguments are accepted and ignored.

Dense math (all nn.Linear, MLN, position encoders, LayerNorm) runs in libfar3d_sm100.so.  Data-dependent bookkeeping
on a few hundred tokens (top-k, gathers, concatenations, the 4x4 pose products of the memory bank) stays in torch
device ops: it is glue between kernels, not arithmetic the roofline sees."""
import math
import os

import torch
import torch.nn as nn

from .. import _lib, ops
from ..compat import BBOX_CODERS, HEADS, TRANSFORMER, build_from_cfg


class MLN(nn.Module):
    """misc.py:153-190 on kernels: gamma/beta GEMMs + one fused (LayerNorm,) scale, shift pass."""

    def __init__(self, c_dim, f_dim=256, use_ln=True):
        super().__init__()
        self.c_dim, self.f_dim, self.use_ln = c_dim, f_dim, use_ln
        self.reduce = nn.Sequential(nn.Linear(c_dim, f_dim), nn.ReLU())
        self.gamma = nn.Linear(f_dim, f_dim)
        self.beta = nn.Linear(f_dim, f_dim)
        if use_ln:
            self.ln = nn.LayerNorm(f_dim, elementwise_affine=False)
        self.init_weight()

    def init_weight(self):
        nn.init.zeros_(self.gamma.weight); nn.init.zeros_(self.beta.weight)
        nn.init.ones_(self.gamma.bias); nn.init.zeros_(self.beta.bias)

    def gamma_beta(self, c):
        h = ops.linear(c.contiguous(), self.reduce[0].weight, self.reduce[0].bias, act=1)
        return (ops.linear(h, self.gamma.weight, self.gamma.bias), ops.linear(h, self.beta.weight, self.beta.bias))

    def forward(self, x, c):
        g, b = self.gamma_beta(c)
        if g.shape != x.shape:                  # `gamma * x + beta` broadcasts in the reference (one code per camera, misc.py:188)
            g, b = g.expand_as(x).contiguous(), b.expand_as(x).contiguous()
        return ops.mln_tokens(x.contiguous(), g, b, self.use_ln)


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def topk_gather(feat, idx):                      # misc.py:13-23
    shape = [1] * feat.dim()
    shape[:2] = idx.shape[:2]
    return torch.gather(feat, 1, idx.view(*shape).repeat(1, 1, *feat.shape[2:]))


def transform_reference_points(pts, pose):       # misc.py:193-202
    p = torch.cat([pts, torch.ones_like(pts[..., :1])], dim=-1)
    return (pose.unsqueeze(1) @ p.unsqueeze(-1)).squeeze(-1)[..., :3]


def denormalize_bbox(b, pc_range=None):          # core/bbox/util.py:25-52
    rot = torch.atan2(b[..., 6:7], b[..., 7:8])
    out = [b[..., 0:1], b[..., 1:2], b[..., 2:3], b[..., 3:4].exp(), b[..., 4:5].exp(), b[..., 5:6].exp(), rot]
    if b.size(-1) > 8:
        out += [b[:, 8:9], b[:, 9:10]]
    return torch.cat(out, dim=-1)


@BBOX_CODERS.register_module()
class NMSFreeCoder:
    """nms_free_coder.py:8-112."""

    def __init__(self, pc_range, voxel_size=None, post_center_range=None, max_num=100, score_threshold=None,
                 num_classes=10):
        self.pc_range, self.voxel_size, self.post_center_range = pc_range, voxel_size, post_center_range
        self.max_num, self.score_threshold, self.num_classes = max_num, score_threshold, num_classes

    fused = True        # far3d_box_decode (one kernel) for CUDA inputs; the torch statement below is what it implements

    def decode_single(self, cls_scores, bbox_preds, bottom_center=False):
        if self.fused and cls_scores.is_cuda and self.max_num <= 512 and bbox_preds.shape[-1] in (8, 10):
            if self.post_center_range is None:
                raise NotImplementedError('post_center_range is required (as in the reference)')
            boxes, scores, labels, _, count = ops.box_decode(cls_scores.contiguous().float(), bbox_preds.contiguous().float(),
                                                             self.max_num, self.post_center_range, self.score_threshold,
                                                             bottom_center)
            n = int(count)                       # the number of boxes inside post_center_range is data dependent (one read)
            return dict(bboxes=boxes[:n], scores=scores[:n], labels=labels[:n].long())
        scores, idx = cls_scores.sigmoid().view(-1).topk(self.max_num)
        labels = idx % self.num_classes
        q = torch.div(idx, self.num_classes, rounding_mode='floor')
        boxes = denormalize_bbox(bbox_preds[q], self.pc_range)
        if self.post_center_range is None:
            raise NotImplementedError('post_center_range is required (as in the reference)')
        r = torch.as_tensor(self.post_center_range, device=scores.device, dtype=boxes.dtype)
        mask = (boxes[..., :3] >= r[:3]).all(1) & (boxes[..., :3] <= r[3:]).all(1)
        if self.score_threshold:
            mask &= scores >= self.score_threshold
        b = boxes[mask]
        if bottom_center:
            b[:, 2] = b[:, 2] - b[:, 5] * 0.5
        return dict(bboxes=b, scores=scores[mask], labels=labels[mask])

    def decode(self, preds_dicts, bottom_center=False):
        cls, box = preds_dicts['all_cls_scores'][-1], preds_dicts['all_bbox_preds'][-1]
        return [self.decode_single(cls[i], box[i], bottom_center) for i in range(cls.size(0))]


def _mlp(x, seq):
    """Run an nn.Sequential of Linear / LayerNorm / ReLU holders through the kernels (ReLU fused into its producer)."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        if isinstance(m, nn.Linear):
            fuse = isinstance(nxt, nn.ReLU)
            x = ops.linear(x, m.weight, m.bias, act=1 if fuse else 0)
            i += 2 if fuse else 1
        elif isinstance(m, nn.LayerNorm):
            fuse = isinstance(nxt, nn.ReLU)
            x = ops.layernorm(x, m.weight, m.bias, m.eps, relu_after=fuse)
            i += 2 if fuse else 1
        elif isinstance(m, nn.ReLU):
            raise RuntimeError('unfused ReLU in kernel MLP')
        else:
            raise TypeError(type(m))
    return x


def _drop_decoder_graphs(module, incompatible_keys):
    module.__dict__.pop('_graphs', None)             # weights changed: captured launches hold stale packed copies


@HEADS.register_module()
class FarHead(nn.Module):
    def __init__(self, num_classes, in_channels=256, stride=16, embed_dims=256, num_query=100, memory_len=1024,
                 topk_proposals=256, num_propagated=256, with_dn=True, with_ego_pos=True, add_query_from_2d=False,
                 depthnet_config={}, train_use_gt_depth=False, val_use_gt_depth=False, add_multi_depth_proposal=False,
                 multi_depth_config={}, return_context_feat=False, return_bbox2d_scores=False, use_offline_2d=False,
                 num_reg_fcs=2, transformer=None, code_weights=None, match_costs=None, bbox_coder=None, code_size=10,
                 normedlinear=False, init_cfg=None, **training_only):
        super().__init__()
        assert not normedlinear and not use_offline_2d and not val_use_gt_depth
        self.num_classes, self.in_channels, self.embed_dims = num_classes, in_channels, embed_dims
        self.cls_out_channels = num_classes
        self.num_query, self.memory_len = num_query, memory_len
        self.topk_proposals, self.num_propagated = topk_proposals, num_propagated
        self.with_ego_pos, self.add_query_from_2d = with_ego_pos, add_query_from_2d
        self.depthnet_config, self.multi_depth_config = dict(depthnet_config), dict(multi_depth_config)
        self.add_multi_depth_proposal = add_multi_depth_proposal
        self.return_context_feat, self.return_bbox2d_scores = return_context_feat, return_bbox2d_scores
        self.code_size, self.num_pred, self.num_reg_fcs = code_size, 6, num_reg_fcs
        cw = (code_weights if code_weights is not None else [1.0] * 8 + [0.2, 0.2])[:code_size]
        mc = match_costs if match_costs is not None else cw
        self.transformer = build_from_cfg(transformer, TRANSFORMER)
        self.code_weights = nn.Parameter(torch.tensor(cw), requires_grad=False)
        self.match_costs = nn.Parameter(torch.tensor(mc), requires_grad=False)
        self.bbox_coder = build_from_cfg(bbox_coder, BBOX_CODERS)
        self.pc_range = nn.Parameter(torch.tensor(self.bbox_coder.pc_range), requires_grad=False)
        cls, reg = [], []
        for _ in range(num_reg_fcs):
            cls += [nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True)]
            reg += [nn.Linear(embed_dims, embed_dims), nn.ReLU()]
        cls.append(nn.Linear(embed_dims, self.cls_out_channels))
        reg.append(nn.Linear(embed_dims, code_size))
        fc_cls, fc_reg = nn.Sequential(*cls), nn.Sequential(*reg)
        self.cls_branches = nn.ModuleList([fc_cls for _ in range(self.num_pred)])      # aliases, farhead.py:248-251
        self.reg_branches = nn.ModuleList([fc_reg for _ in range(self.num_pred)])
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
        # captured decoder graphs bake weight addresses: drop them whenever weights are reloaded or moved
        self.register_load_state_dict_post_hook(_drop_decoder_graphs)

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop('_graphs', None)
        self.__dict__.pop('_feat_buf', None)
        return super()._apply(fn, *args, **kwargs)

    def init_weights(self):                       # farhead.py:432-444
        nn.init.uniform_(self.reference_points.weight.data, 0, 1)
        if self.num_propagated > 0:
            nn.init.uniform_(self.pseudo_reference_points.weight.data, 0, 1)
            self.pseudo_reference_points.weight.requires_grad = False
        self.transformer.init_weights()
        nn.init.constant_(self.cls_branches[0][-1].bias, float(-math.log((1 - 0.01) / 0.01)))

    # ------------------------------------------------------------------ memory bank (farhead.py:446-508)
    def reset_memory(self):
        self.memory_embedding = self.memory_reference_point = self.memory_timestamp = None
        self.memory_egopose = self.memory_velo = None

    MEMORY_KEYS = ('memory_embedding', 'memory_reference_point', 'memory_timestamp', 'memory_egopose', 'memory_velo')

    def export_memory(self):
        """the temporal memory bank of the camera-rig stream served last (the tensors themselves: every frame re-binds the
        attributes to new tensors, farhead.py:479-508, so holding these is a snapshot)"""
        return {k: getattr(self, k) for k in self.MEMORY_KEYS}

    def import_memory(self, state):
        """make `state` (from export_memory; None = a stream that has not been seen yet) the live bank"""
        for k in self.MEMORY_KEYS:
            setattr(self, k, None if state is None else state[k])

    # far3d_memory_pre_update / far3d_memory_post_update instead of ~35 torch launches per frame (FAR3D_MEMORY_KERNELS=0: torch glue)
    memory_kernels = os.environ.get('FAR3D_MEMORY_KERNELS', '1') != '0'

    def _bank(self):
        """the live bank as the kernels take it (one stream: B = 1), contiguous, timestamps fp64"""
        return dict(emb=self.memory_embedding[0].contiguous(), ref=self.memory_reference_point[0].contiguous(),
                    ts=self.memory_timestamp[0, :, 0].double().contiguous(), pose=self.memory_egopose[0].contiguous(),
                    velo=self.memory_velo[0].contiguous())

    def _set_bank(self, b):
        self.memory_embedding, self.memory_reference_point = b['emb'].unsqueeze(0), b['ref'].unsqueeze(0)
        self.memory_timestamp, self.memory_egopose, self.memory_velo = b['ts'].view(1, -1, 1), b['pose'].unsqueeze(0), b['velo'].unsqueeze(0)

    def _use_memory_kernels(self, x):
        return self.memory_kernels and x.is_cuda and x.size(0) == 1

    def pre_update_memory(self, data):
        x = data['prev_exists']
        B = x.size(0)
        pr = self.pc_range
        if self.memory_embedding is not None and self._use_memory_kernels(x):
            k = self.num_propagated
            pseudo = (self.pseudo_reference_points.weight * (pr[3:6] - pr[0:3]) + pr[0:3]).contiguous() if k > 0 else None
            self._set_bank(ops.memory_pre_update(self._bank(), self.memory_len, k, x.float().contiguous(),
                                                 data['ego_pose_inv'][0].float().contiguous(),
                                                 data['timestamp'].double().contiguous(), pseudo))
            return
        if self.memory_embedding is None:
            self.memory_embedding = x.new_zeros(B, self.memory_len, self.embed_dims)
            self.memory_reference_point = x.new_zeros(B, self.memory_len, 3)
            self.memory_timestamp = x.new_zeros(B, self.memory_len, 1)
            self.memory_egopose = x.new_zeros(B, self.memory_len, 4, 4)
            self.memory_velo = x.new_zeros(B, self.memory_len, 2)
        else:
            n = self.memory_len
            self.memory_timestamp = (self.memory_timestamp + data['timestamp'].unsqueeze(-1).unsqueeze(-1))[:, :n] * x.view(-1, 1, 1)
            self.memory_egopose = (data['ego_pose_inv'].unsqueeze(1) @ self.memory_egopose)[:, :n] * x.view(-1, 1, 1, 1)
            self.memory_reference_point = transform_reference_points(self.memory_reference_point,
                                                                     data['ego_pose_inv'])[:, :n] * x.view(-1, 1, 1)
            self.memory_embedding = self.memory_embedding[:, :n] * x.view(-1, 1, 1)
            self.memory_velo = self.memory_velo[:, :n] * x.view(-1, 1, 1)
        if self.num_propagated > 0:
            k = self.num_propagated
            pseudo = self.pseudo_reference_points.weight * (pr[3:6] - pr[0:3]) + pr[0:3]
            self.memory_reference_point = self.memory_reference_point.clone()
            self.memory_egopose = self.memory_egopose.clone()
            self.memory_reference_point[:, :k] += (1 - x).view(B, 1, 1) * pseudo
            self.memory_egopose[:, :k] += (1 - x).view(B, 1, 1, 1) * torch.eye(4, device=x.device)

    def post_update_memory(self, data, rec_ego_pose, all_cls_scores, all_bbox_preds, outs_dec):
        if self._use_memory_kernels(all_cls_scores) and all_cls_scores.shape[2] <= 4096:
            # (rec_ego_pose is the identity for every query: farhead.py:310)
            new, idx = ops.memory_post_update(all_cls_scores[-1][0].contiguous(), all_bbox_preds[-1][0].contiguous(),
                                              outs_dec[-1][0].contiguous(), self.topk_proposals,
                                              data['ego_pose'][0].float().contiguous(), data['timestamp'].double().contiguous(),
                                              self._bank())
            self.last_topk_indexes = idx.long().view(1, -1, 1)
            self._set_bank(new)
            return
        rec_score = all_cls_scores[-1].sigmoid().topk(1, dim=-1).values[..., 0:1]
        _, idx = torch.topk(rec_score, self.topk_proposals, dim=1)
        rec_ts = topk_gather(torch.zeros_like(rec_score, dtype=torch.float64), idx)
        rec_ref = topk_gather(all_bbox_preds[..., :3][-1], idx)
        rec_memory = topk_gather(outs_dec[-1], idx)
        rec_pose = topk_gather(rec_ego_pose, idx)
        rec_velo = topk_gather(all_bbox_preds[..., -2:][-1], idx)
        self.last_topk_indexes = idx
        self.memory_embedding = torch.cat([rec_memory, self.memory_embedding], dim=1)
        self.memory_timestamp = torch.cat([rec_ts, self.memory_timestamp], dim=1) - data['timestamp'].unsqueeze(-1).unsqueeze(-1)
        self.memory_egopose = data['ego_pose'].unsqueeze(1) @ torch.cat([rec_pose, self.memory_egopose], dim=1)
        self.memory_reference_point = transform_reference_points(
            torch.cat([rec_ref, self.memory_reference_point], dim=1), data['ego_pose'])
        self.memory_velo = torch.cat([rec_velo, self.memory_velo], dim=1)

    def _pos3d(self, ref):
        return _mlp(ops.pos2posemb3d(ref.contiguous().float()), self.query_embedding)

    def temporal_alignment(self, query_pos, tgt, reference_points):        # farhead.py:284-313
        B, Q = query_pos.shape[:2]
        pr = self.pc_range
        dev = query_pos.device
        temp_ref = (self.memory_reference_point - pr[:3]) / (pr[3:6] - pr[0:3])
        temp_pos = self._pos3d(temp_ref)
        temp_memory = self.memory_embedding
        eye = torch.eye(4, device=dev)
        if self.with_ego_pos:
            rec_motion = torch.cat([reference_points.new_zeros(B, Q, 3), eye[:3, :].flatten().expand(B, Q, 12)], dim=-1)
            rec_pe = ops.nerf_posenc(rec_motion.contiguous())
            tgt = self.ego_pose_memory(tgt, rec_pe)
            query_pos = self.ego_pose_pe(query_pos, rec_pe)
            mem_motion = torch.cat([self.memory_velo, self.memory_timestamp, self.memory_egopose[..., :3, :].flatten(-2)],
                                   dim=-1).float()
            mem_pe = ops.nerf_posenc(mem_motion.contiguous())
            temp_pos = self.ego_pose_pe(temp_pos, mem_pe)
            temp_memory = self.ego_pose_memory(temp_memory, mem_pe)
        te = self.time_embedding
        q_t = ops.pos2posemb1d(reference_points.new_zeros(B, Q, 1))
        query_pos = ops.layernorm(ops.linear(q_t, te[0].weight, te[0].bias), te[1].weight, te[1].bias, te[1].eps) + query_pos
        m_t = ops.pos2posemb1d(self.memory_timestamp.float().contiguous())
        temp_pos = ops.layernorm(ops.linear(m_t, te[0].weight, te[0].bias), te[1].weight, te[1].bias, te[1].eps) + temp_pos
        if self.num_propagated > 0:
            k = self.num_propagated
            tgt = torch.cat([tgt, temp_memory[:, :k]], dim=1)
            query_pos = torch.cat([query_pos, temp_pos[:, :k]], dim=1)
            reference_points = torch.cat([reference_points, temp_ref[:, :k]], dim=1)
            temp_memory = temp_memory[:, k:].contiguous()
            temp_pos = temp_pos[:, k:].contiguous()
        rec_ego_pose = eye.view(1, 1, 4, 4).repeat(B, tgt.shape[1], 1, 1)
        return tgt, query_pos, reference_points.contiguous(), temp_memory, temp_pos, rec_ego_pose

    # ------------------------------------------------------------------ feature prep (farhead.py:553-567)
    def flatten_features(self, mlvl_feats, data):
        intr = data['intrinsics'] / 1e3
        extr = data['extrinsics'][..., :3, :]
        mln_in = torch.cat([intr[..., 0, 0:1], intr[..., 1, 1:2], extr.flatten(-2)], dim=-1).flatten(0, 1)   # [BN,14]
        gamma, beta = self.spatial_alignment.gamma_beta(mln_in)
        shapes = [tuple(f.shape[-2:]) for f in mlvl_feats]
        starts, s = [], 0
        for h, w in shapes:
            starts.append(s)
            s += h * w
        B, N, C = mlvl_feats[0].shape[:3]
        # persistent destination: a stable address lets the captured decoder graph read it in place
        out = getattr(self, '_feat_buf', None)
        if out is None or tuple(out.shape) != (B * N, s, C) or out.device != mlvl_feats[0].device:
            out = self._feat_buf = torch.empty(B * N, s, C, device=mlvl_feats[0].device)
        for f, (h, w), st in zip(mlvl_feats, shapes, starts):
            f = f.flatten(0, 1) if f.dim() == 5 else f
            if f.stride(1) == 1 and f.permute(0, 2, 3, 1).is_contiguous():      # channels-last (our FPN)
                ops.mln_flatten(f.permute(0, 2, 3, 1).reshape(B * N, h * w, C), gamma, beta, out, st, True)
            else:
                ops.mln_flatten(f.contiguous().view(B * N, C, h * w), gamma, beta, out, st, False)
        dev = out.device
        spatial = torch.as_tensor(shapes, dtype=torch.long, device=dev)
        level_start = torch.as_tensor(starts, dtype=torch.long, device=dev)
        # host copies ride along so the aggregation op needs no device->host read
        self._levels_host = (tuple(shapes), tuple(starts))
        return out, spatial, level_start

    def _convert_bin_depth_to_specific(self, idx, inverse=False):        # farhead.py:521-531
        dmin, dmax, nb = [self.depthnet_config.get(k) for k in ('depth_min', 'depth_max', 'num_depth_bins')]
        bs = 2 * (dmax - dmin) / (nb * (1 + nb))
        if not inverse:
            return dmin + bs / 8 * (torch.square(idx / 0.5 + 1) - 1)
        return (-0.5 + 0.5 * torch.sqrt(1 + 8 * (idx - dmin) / bs)).type(torch.int64)

    @torch.no_grad()
    def build_query2d_proposal(self, bbox_list, pred_depth, data, bn, padHW, context2d_feat=None, bbox2d_scores=None):
        """farhead.py:710-827 for depth-logit input (multi_depth_config topk set) and B == 1. Data-dependent sizes:
        torch device ops (this is row a6 / f1 of SURVEY.md section 8, outside the kernel path for now)."""
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
            c = (bb[:, :2] / down).round().long().clamp(min=0)
            c[:, 0] = c[:, 0].clamp(max=w_max - 1)
            c[:, 1] = c[:, 1].clamp(max=h_max - 1)
            depths.append(dm[(c[:, 1] * (pad_w / down) + c[:, 0]).long()])
        depths = torch.cat(depths, dim=0)
        topk = self.multi_depth_config.get('topk', -1)
        assert topk != -1, 'far3d_b200 implements the depth-logit proposal path (far3d.py:93)'
        ok = None
        if self.add_multi_depth_proposal:
            rmin = self.__dict__.get('_rmin_bin')          # config constant: computed once (no per-frame host round trip)
            if rmin is None:
                rmin = self._convert_bin_depth_to_specific(torch.tensor([float(self.multi_depth_config.get('range_min', -1))]),
                                                           inverse=True).item()
                self.__dict__['_rmin_bin'] = rmin
            tv, ti = torch.topk(depths, topk, dim=1)
            ok = ti[:, 0] >= rmin
            boxes = torch.cat([boxes, boxes.repeat(topk - 1, 1)[ok.repeat(topk - 1)]], dim=0)
            depths = torch.cat([ti[:, 0:1], ti[:, 1:][ok].transpose(1, 0).flatten().unsqueeze(-1)], dim=0)
            if context2d_feat is not None:
                context2d_feat = torch.cat([context2d_feat, context2d_feat.repeat(topk - 1, 1)[ok.repeat(topk - 1)]], dim=0)
            if bbox2d_scores is not None:
                thr = torch.tensor([0.1], device=bbox2d_scores.device)
                lo = torch.log(bbox2d_scores / (1 - bbox2d_scores)) - torch.log(thr / (1 - thr))
                tv = tv / tv[:, 0:1]
                ds = torch.cat([tv[:, 0:1], tv[:, 1:][ok].transpose(1, 0).flatten().unsqueeze(-1)], dim=0)
                lo = torch.cat([lo, lo[ok].repeat(topk - 1, 1)], dim=0) * ds
                context2d_feat = torch.cat([context2d_feat, lo], dim=-1) if context2d_feat is not None \
                    else lo.repeat(1, self.in_channels)
        else:
            depths = torch.argmax(depths, dim=-1, keepdim=True)
        depths = self._convert_bin_depth_to_specific(depths)
        coords = torch.cat([boxes[:, :2], depths, torch.ones_like(depths)], dim=1)
        coords[..., :2] = coords[..., :2] * torch.maximum(coords[..., 2:3], torch.ones_like(coords[..., 2:3]) * 1e-5)
        i2l = data['lidar2img'].inverse().view(B * N, 1, 4, 4)
        i2l = torch.cat([i2l[k].repeat(n, 1, 1) for k, n in enumerate(nums)], dim=0)
        if ok is not None:
            i2l = torch.cat([i2l, i2l.repeat(topk - 1, 1, 1)[ok.repeat(topk - 1)]], dim=0)
        c3 = torch.matmul(i2l, coords.unsqueeze(-1)).squeeze(-1)[..., :3]
        pr = self.pc_range
        c3 = (c3 - pr[:3]) / (pr[3:6] - pr[:3])
        if B != 1:
            raise NotImplementedError
        return c3.unsqueeze(0), (context2d_feat.unsqueeze(0) if context2d_feat is not None else None)

    # ------------------------------------------------------------------ adaptive queries on the device
    proposal_kernels = True    # far3d_roi_select / far3d_query2d_lift instead of the boolean-gather torch glue
    proposal_cap = 2048        # capacity of the adaptive-query block (cfg-4 at the 0.05 test head: ~1900)
    proposal_bucket = 64       # the block is padded to a multiple of this: the decoder graph is keyed by the PADDED count

    def _proposals_device(self, outs_roi, feat_flatten, data, img_metas):
        """farhead.py:710-827 + :585-602 without boolean gathers: fixed-capacity kernels, then ONE small device->host read (the
        query count picks the padded size / the captured decoder graph).  Returns (ref2d [1,Mpad,3] | None, ctx [1,Mpad,C+1] |
        None, real count Mq, padded count Mpad)."""
        sel, lg, nb = outs_roi['_sel'], outs_roi['_depth_logit_nhwc'], outs_roi['_depth_bins']
        assert data['lidar2img'].shape[0] == 1, 'one sample per frame (farhead.py:813-816 raises for B > 1 too)'
        N = sel['N']
        S = feat_flatten.shape[1]
        assert sel['S2'] == S, 'the 2D head and the feature pyramid must cover the same levels (farhead.py:588 indexes feat_flatten with the 2D mask)'
        pad_h = img_metas[0]['pad_shape'][0][0]
        down = int(pad_h / lg.shape[1])
        cfg = self.depthnet_config
        dmin, dmax, nbins = float(cfg.get('depth_min')), float(cfg.get('depth_max')), int(cfg.get('num_depth_bins'))
        bs = 2 * (dmax - dmin) / (nbins * (1 + nbins))
        topk = int(self.multi_depth_config.get('topk', -1)) if self.add_multi_depth_proposal else 1
        assert 1 <= topk <= 4, 'far3d_b200 implements the depth-logit proposal path (far3d.py:93)'
        rmin = self.__dict__.get('_rmin_bin')
        if rmin is None:
            rmin = self.__dict__['_rmin_bin'] = int(-0.5 + 0.5 * math.sqrt(1 + 8 * (float(self.multi_depth_config.get('range_min', -1)) - dmin) / bs))
        cap = self.proposal_cap
        ref_all, src_row, score_feat, meta = ops.query2d_lift(sel, lg, nb, down, topk, rmin, dmin, bs, math.log(0.1 / 0.9),
                                                              data['lidar2img'][0].contiguous(), self.pc_range, S, cap)
        ctx_all = None
        if self.return_context_feat and self.return_bbox2d_scores:
            ctx_all = ops.ctx_gather(feat_flatten, src_row, score_feat, cap)
        st = self.__dict__.get('_prop_host')
        if st is None:
            st = self.__dict__['_prop_host'] = dict(pin=torch.empty(4 + 16, dtype=torch.int32, pin_memory=True),
                                                   skip_pin=torch.empty(2, dtype=torch.int32, pin_memory=True),
                                                   skip=torch.zeros(2, dtype=torch.int32, device=feat_flatten.device),
                                                   ev=torch.cuda.Event())
        st['pin'][:4].copy_(meta, non_blocking=True)
        st['pin'][4:4 + N].copy_(sel['counts'], non_blocking=True)
        st['ev'].record()
        st['ev'].synchronize()                           # the frame's only device->host dependency before the decoder
        mq, over = int(st['pin'][0]), int(st['pin'][3])
        counts = [min(int(c), sel['cap']) for c in st['pin'][4:4 + N].tolist()]
        outs_roi['bbox_list'] = [sel['box'][i, :c] for i, c in enumerate(counts)]
        outs_roi['bbox2d_scores'] = torch.cat([sel['score'][i, :c] for i, c in enumerate(counts)]).reshape(-1, 1)
        if over:
            raise RuntimeError(f'2D proposal capacity exceeded ({st["pin"][4:4 + N].tolist()} peaks per camera, cap {sel["cap"]} / '
                               f'{cap} queries): raise YOLOXHeadCustom.select_cap / FarHead.proposal_cap')
        if mq == 0:
            return None, None, 0, 0
        b = self.proposal_bucket
        mpad = min(cap, -(-mq // b) * b)
        st['skip_pin'][0], st['skip_pin'][1] = self.num_query + mq, mpad - mq
        st['skip'].copy_(st['skip_pin'], non_blocking=True)
        ctx = ctx_all[:mpad].unsqueeze(0) if ctx_all is not None else None
        return ref_all[:mpad].unsqueeze(0), ctx, mq, mpad

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, img_metas, outs_roi=None, **data):
        if self.training:
            raise RuntimeError('far3d_b200 FarHead implements the inference forward only; call .eval()')
        self.pre_update_memory(data)
        pre = data.get('_feat_flatten')
        if pre is not None:             # camera-sharded caller (parallel.CameraShardedFar3D): maps already MLN'd, flattened, gathered
            feat_flatten, spatial_flatten, level_start_index = pre
            B, N = data['lidar2img'].shape[:2]
        else:
            mlvl_feats = data['img_feats']
            B, N = mlvl_feats[0].shape[:2]
            feat_flatten, spatial_flatten, level_start_index = self.flatten_features(mlvl_feats, data)
        reference_points = self.reference_points.weight.unsqueeze(0).repeat(B, 1, 1)
        query_pos = self._pos3d(reference_points)
        ref2d = ctx = None
        npro = mq = 0
        if self.add_query_from_2d and outs_roi is not None and outs_roi.get('_sel') is not None and self.proposal_kernels:
            ref2d, ctx, mq, npro = self._proposals_device(outs_roi, feat_flatten, data, img_metas)
            if ref2d is not None:
                query_pos = torch.cat([query_pos, self._pos3d(ref2d)], dim=1)
                reference_points = torch.cat([reference_points, ref2d], dim=1)
        elif self.add_query_from_2d and outs_roi is not None:
            scores = outs_roi['bbox2d_scores'] if self.return_bbox2d_scores else None
            ctx2d = None
            if self.return_context_feat:
                vi = outs_roi['valid_indices']
                ctx2d = feat_flatten[vi.repeat(1, 1, feat_flatten.shape[-1])].reshape(-1, feat_flatten.shape[-1])
            padHW = img_metas[0]['pad_shape'][0][:2]
            ref2d, ctx = self.build_query2d_proposal(outs_roi['bbox_list'], outs_roi['pred_depth'].permute(0, 2, 3, 1), data,
                                                     (B, N), padHW, ctx2d, scores)
            if ref2d is not None:
                npro = mq = ref2d.shape[1]
                query_pos = torch.cat([query_pos, self._pos3d(ref2d)], dim=1)
                reference_points = torch.cat([reference_points, ref2d], dim=1)
        tgt = torch.zeros_like(query_pos)
        if ctx is not None:
            tgt[:, -npro:, :] = _mlp(ctx.contiguous(), self.context_embed)
        tgt, query_pos, reference_points, temp_memory, temp_pos, rec_ego_pose = \
            self.temporal_alignment(query_pos, tgt, reference_points)
        padded = npro > mq                     # rows [num_query + mq, num_query + npro) are padding of the bucketed count
        dev_path = npro > 0 and self.proposal_kernels and outs_roi is not None and outs_roi.get('_sel') is not None
        ops.MHA_KEY_SKIP = self.__dict__['_prop_host']['skip'] if dev_path else None
        try:
            outs_dec, all_cls, all_box = self._decode(tgt, query_pos, feat_flatten, temp_memory, temp_pos, reference_points,
                                                      data['lidar2img'], img_metas)
        finally:
            ops.MHA_KEY_SKIP = None
        if padded:                             # drop the padding rows: everything downstream sees the reference's shapes
            a, b_ = self.num_query + mq, self.num_query + npro
            keep = lambda t, dim: torch.cat([t.narrow(dim, 0, a), t.narrow(dim, b_, t.shape[dim] - b_)], dim=dim)
            outs_dec, all_cls, all_box = keep(outs_dec, 2), keep(all_cls, 2), keep(all_box, 2)
            reference_points, rec_ego_pose = keep(reference_points, 1), keep(rec_ego_pose, 1)
            ref2d = ref2d[:, :mq]
        ref_logit = inverse_sigmoid(reference_points.clone())
        all_box[..., 0:3] = (all_box[..., 0:3] + ref_logit[None, ..., 0:3]).sigmoid()
        pr = self.pc_range
        all_box[..., 0:3] = all_box[..., 0:3] * (pr[3:6] - pr[0:3]) + pr[0:3]
        self.post_update_memory(data, rec_ego_pose, all_cls, all_box, outs_dec)
        return dict(all_cls_scores=all_cls, all_bbox_preds=all_box, dn_mask_dict=None, reference_points2d=ref2d,
                    outs_dec=outs_dec, feat_flatten=feat_flatten, spatial_flatten=spatial_flatten,
                    level_start_index=level_start_index)

    # ------------------------------------------------------------------ decoder + branches, optionally as a CUDA graph
    use_cuda_graph = True          # the ~200 short decoder launches are host-bound otherwise (r1 section timing)

    def _decode_eager(self, tgt, query_pos, feat_flatten, temp_memory, temp_pos, reference_points, lidar2img, img_metas):
        # host-side level tables go to the decoder: no device->host read per layer
        outs_dec = self.transformer(tgt, query_pos, feat_flatten, self._levels_host[0], self._levels_host[1], temp_memory,
                                    temp_pos, None, reference_points, self.pc_range, dict(lidar2img=lidar2img), img_metas)
        outs_dec = torch.nan_to_num(outs_dec)
        L, Bq, Q, E = outs_dec.shape
        # the 6 branch modules are one shared module (farhead.py:248-251): run all layers' tokens in one GEMM chain
        flat = outs_dec.reshape(L * Bq * Q, E)
        all_cls = _mlp(flat, self.cls_branches[0]).view(L, Bq, Q, -1)
        all_box = _mlp(flat, self.reg_branches[0]).view(L, Bq, Q, -1)
        return outs_dec, all_cls, all_box

    def _decode(self, tgt, query_pos, feat_flatten, temp_memory, temp_pos, reference_points, lidar2img, img_metas):
        args = (tgt, query_pos, temp_memory, temp_pos, reference_points, lidar2img)
        if not self.use_cuda_graph or ops.PROFILE is not None or not tgt.is_cuda:
            return self._decode_eager(tgt, query_pos, feat_flatten, temp_memory, temp_pos, reference_points, lidar2img, img_metas)
        pad = tuple(img_metas[0]['pad_shape'][0][:2])
        key = (tuple(tgt.shape), tuple(temp_memory.shape), tuple(feat_flatten.shape), feat_flatten.data_ptr(), pad,
               self._levels_host, ops.LINEAR_MODE, None if ops.MHA_KEY_SKIP is None else ops.MHA_KEY_SKIP.data_ptr())
        cache = self.__dict__.setdefault('_graphs', {})
        ent = cache.get(key)
        if ent is None:
            if len(cache) >= 8:                       # adaptive-query count varies frame to frame: keep a small LRU
                cache.pop(next(iter(cache)))
            static = [a.detach().clone().contiguous() for a in args]
            run = lambda: self._decode_eager(static[0], static[1], feat_flatten, static[2], static[3], static[4], static[5], img_metas)
            run()                                     # warm-up outside capture: fills the packed-weight caches
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g):
                outs = run()
            ent = cache[key] = (g, static, outs, _lib.launch_count() - n0)
        else:
            cache[key] = cache.pop(key)               # refresh LRU position
        g, static, outs, nlaunch = ent
        for s_, a in zip(static, args):
            s_.copy_(a)
        g.replay()
        _lib.load().far3d_add_launches(nlaunch)       # the replay re-launches the captured kernels
        return tuple(o.clone() for o in outs)

    def get_bboxes(self, preds_dicts, img_metas=None, rescale=False):      # farhead.py:1224-1245
        ret = []
        for i, p in enumerate(self.bbox_coder.decode(preds_dicts, bottom_center=True)):
            b = p['bboxes']
            box_type = (img_metas[i].get('box_type_3d') if img_metas is not None else None)
            ret.append([box_type(b, b.size(-1)) if box_type is not None else b, p['scores'], p['labels']])
        return ret
