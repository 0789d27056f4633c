# Model section of Far3D's Argoverse-2 configuration, restated for far3d_b200 (inference only).
# It is value-for-value the `model` dict of the reference's projects/configs/far3d.py:38-147 (tests/test_cpu.py asserts
# equality against that file when the reference tree is present); datasets, pipelines, optimiser and schedules of the
# reference config are out of scope.  The reference file itself also loads unchanged through
# far3d_b200.compat.Config.fromfile.
plugin = True
plugin_dir = 'far3d_b200/plugin/'

point_cloud_range = [-152.4, -152.4, -5.0, 152.4, 152.4, 5.0]
voxel_size = [0.2, 0.2, 8]
# projects/configs/far3d.py:13-20
img_norm_cfg = dict(mean=[103.530, 116.280, 123.675], std=[57.375, 57.120, 58.395], to_rgb=False)
class_names = ['ARTICULATED_BUS', 'BICYCLE', 'BICYCLIST', 'BOLLARD', 'BOX_TRUCK', 'BUS',
               'CONSTRUCTION_BARREL', 'CONSTRUCTION_CONE', 'DOG', 'LARGE_VEHICLE',
               'MESSAGE_BOARD_TRAILER', 'MOBILE_PEDESTRIAN_CROSSING_SIGN', 'MOTORCYCLE',
               'MOTORCYCLIST', 'PEDESTRIAN', 'REGULAR_VEHICLE', 'SCHOOL_BUS', 'SIGN',
               'STOP_SIGN', 'STROLLER', 'TRUCK', 'TRUCK_CAB', 'VEHICULAR_TRAILER',
               'WHEELCHAIR', 'WHEELED_DEVICE', 'WHEELED_RIDER']
num_classes = 26
embed_dims = 256
depthnet_config = {'type': 0, 'hidden_dim': 256, 'num_depth_bins': 50, 'depth_min': 1e-1, 'depth_max': 110, 'stride': 8}

_self_attn = dict(type='MultiheadAttention', embed_dims=256, num_heads=8, dropout=0.1)
_cross_attn = dict(type='DeformableFeatureAggregationCuda', embed_dims=256, num_groups=8, num_levels=4, num_cams=7,
                   dropout=0.1, num_pts=13, bias=2.)
_decoder_layer = dict(
    type='Detr3DTemporalDecoderLayer', batch_first=True, attn_cfgs=[_self_attn, _cross_attn],
    feedforward_channels=2048, ffn_dropout=0.1, with_cp=True,
    operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))

model = dict(
    type='Far3D',
    use_grid_mask=True,
    stride=[8, 16, 32, 64],
    position_level=[0, 1, 2, 3],
    img_backbone=dict(type='VoVNet', spec_name='V-99-eSE', norm_eval=True, frozen_stages=-1, input_ch=3,
                      out_features=('stage2', 'stage3', 'stage4', 'stage5',)),
    img_neck=dict(type='FPN', start_level=1, add_extra_convs='on_output', relu_before_extra_convs=True,
                  in_channels=[256, 512, 768, 1024], out_channels=256, num_outs=4),
    img_roi_head=dict(
        type='YOLOXHeadCustom', num_classes=26, in_channels=256, strides=[8, 16, 32, 64],
        train_cfg=dict(assigner=dict(type='SimOTAAssigner', center_radius=2.5)),
        test_cfg=dict(score_thr=0.01, nms=dict(type='nms', iou_threshold=0.65)),
        pred_with_depth=True, depthnet_config=depthnet_config, reg_depth_level='p3', pred_depth_var=False,
        loss_depth2d=dict(type='L1Loss', loss_weight=1.0), sample_with_score=True, threshold_score=0.1,
        topk_proposal=None, return_context_feat=True),
    pts_bbox_head=dict(
        type='FarHead', num_classes=26, in_channels=256, num_query=644, memory_len=1024, topk_proposals=256,
        num_propagated=256, scalar=10, noise_scale=1.0, dn_weight=1.0, split=0.75, offset=0.5, offset_p=0.0,
        num_smp_per_gt=3, with_dn=True, with_ego_pos=True, add_query_from_2d=True, pred_box_var=False,
        depthnet_config=depthnet_config, train_use_gt_depth=True, add_multi_depth_proposal=True,
        multi_depth_config={'topk': 1, 'range_min': 30, }, return_bbox2d_scores=True, return_context_feat=True,
        code_size=8, code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0],
        transformer=dict(type='Detr3DTransformer',
                         decoder=dict(type='Detr3DTransformerDecoder', embed_dims=256, num_layers=6,
                                      transformerlayers=_decoder_layer)),
        bbox_coder=dict(type='NMSFreeCoder', post_center_range=point_cloud_range, pc_range=point_cloud_range, max_num=300,
                        voxel_size=voxel_size, num_classes=26),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
        loss_bbox=dict(type='L1Loss', loss_weight=0.25),
        loss_iou=dict(type='GIoULoss', loss_weight=0.0), ),
    train_cfg=dict(pts=dict(
        grid_size=[512, 512, 1], voxel_size=voxel_size, point_cloud_range=point_cloud_range, out_size_factor=4,
        assigner=dict(type='HungarianAssigner3D', cls_cost=dict(type='FocalLossCost', weight=2.0),
                      reg_cost=dict(type='BBox3DL1Cost', weight=0.25), iou_cost=dict(type='IoUCost', weight=0.0),
                      pc_range=point_cloud_range), )))
