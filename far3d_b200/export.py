"""What leaves the per-frame path: detections as the reference hands them to the Argoverse 2 evaluation / submission tools.

`results_to_av2(results, frame_infos, class_names)` restates `Argoverse2Dataset.format_results` + `box_to_av2`
(projects/mmdet3d_plugin/datasets/argoverse2_dataset.py:267-346) and `yaw_to_quat` (datasets/av2_utils.py:241-282) for the
output of `Far3D.simple_test` / `Far3DPipeline.infer`: per frame a dict `pts_bbox = {boxes_3d (K,7) [x, y, z_bottom, w, l, h,
yaw], scores_3d (K,), labels_3d (K,)}` (mmdet3d `bbox3d2result`, far3d.py:268-277).  Host-side Python like the reference's
(pandas / pyarrow); the detections arrive in one 12 KB device-to-host copy per frame (`Far3DPipeline._to_host`)."""
import numpy as np
import torch

# argoverse2_dataset.py:15-17
LABEL_ATTR = ('tx_m', 'ty_m', 'tz_m', 'length_m', 'width_m', 'height_m', 'qw', 'qx', 'qy', 'qz')


def yaw_to_quat(yaw_rad):
    """scalar-first quaternions (w, x, y, z) of rotations about z: av2_utils.py:241-282 with roll = pitch = 0, evaluated with
    the same products so the result is identical, signed zeros included"""
    z = torch.as_tensor(yaw_rad)
    zero = torch.zeros_like(z)
    cy, sy = torch.cos(z * 0.5), torch.sin(z * 0.5)
    cp, sp, cr, sr = torch.cos(zero), torch.sin(zero), torch.cos(zero), torch.sin(zero)
    return torch.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                        cr * cp * sy - sr * sp * cy], dim=-1)


def box_to_av2(boxes_3d):
    """(K, 7) boxes [x, y, z_bottom, w, l, h, yaw] -> (K, 10) [gravity centre, tensor[:, 3:6], quaternion]
    (argoverse2_dataset.py:339-346; `gravity_center` of mmdet3d's LiDARInstance3DBoxes = bottom centre + h / 2 on z)"""
    t = torch.as_tensor(getattr(boxes_3d, 'tensor', boxes_3d)).detach().cpu().float()
    centre = t[:, :3].clone()
    centre[:, 2] = t[:, 2] + t[:, 5] * 0.5
    return torch.cat([centre, t[:, [3, 4, 5]], yaw_to_quat(t[:, 6])], dim=1)


def results_to_av2(results, frame_infos, class_names, feather_path=None):
    """results: list over frames of {'pts_bbox': {...}} (or the inner dict); frame_infos: list of dicts with `scene_id` and
    `lidar_timestamp_ns` (the AV2 info .pkl schema, argoverse2_dataset_t.py:162-240).  Returns the detections DataFrame indexed
    by (log_id, timestamp_ns) as `format_results` does; `feather_path` also writes the score-sorted table the AV2 tools read."""
    import pandas as pd
    assert len(results) == len(frame_infos), (len(results), len(frame_infos))
    frames = []
    for out, info in zip(results, frame_infos):
        out = out.get('pts_bbox', out)
        labels = torch.as_tensor(out['labels_3d']).cpu().numpy().tolist()
        df = pd.DataFrame(box_to_av2(out['boxes_3d']).numpy(), columns=list(LABEL_ATTR))
        df['score'] = torch.as_tensor(out['scores_3d']).detach().cpu().numpy()
        df['log_id'] = info['scene_id']
        df['timestamp_ns'] = int(info['lidar_timestamp_ns'])
        df['category'] = [class_names[i].upper() for i in labels]
        frames.append(df)
    dts = pd.concat(frames).set_index(['log_id', 'timestamp_ns']).sort_index()
    dts = dts.sort_values('score', ascending=False).reset_index()
    if feather_path is not None:
        if not str(feather_path).endswith('.feather'):
            feather_path = f'{feather_path}.feather'
        dts.to_feather(feather_path)
    return dts.set_index(['log_id', 'timestamp_ns']).sort_index()
