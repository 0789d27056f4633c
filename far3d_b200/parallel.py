"""Multi-GPU plan for the per-frame path: one process per GPU, `torch.distributed` (NCCL over
NVLink 5 / NVSwitch on the box, gloo in CPU tests).

Two ways the path shards, both used by bench.py:
  * streams / frames are independent (each owns a memory bank) - the reference's own test-time sharding
    (datasets/samplers/distributed_sampler.py:41-44): ranks run whole frames, no data-path collective ("replicas");
  * cameras are independent through backbone -> FPN -> MLN/flatten (everything before farhead.py:565); the decoder is
    cross-camera.  `CameraShardedFar3D` runs the image branch on a contiguous camera slice per rank and issues ONE
    all-gather of the flattened channels-last maps before the (replicated) decoder.
"""
import torch
import torch.distributed as dist


def shard_cameras(num_cams, world_size):
    """Contiguous [begin, end) camera ranges per rank; ranks beyond the camera count get empty ranges."""
    per = -(-num_cams // world_size)
    return [(min(r * per, num_cams), min((r + 1) * per, num_cams)) for r in range(world_size)]


def all_gather_cameras(local, num_cams, group=None):
    """local [n_local, S, C] (this rank's cameras, may be empty) -> [num_cams, S, C] on every rank.
    One collective; ragged shards are padded to the largest shard so a single fixed-size all-gather suffices."""
    world = dist.get_world_size(group)
    plan = shard_cameras(num_cams, world)
    per = max(b - a for a, b in plan)
    S, C = local.shape[1:]
    buf = local
    if local.shape[0] != per:
        buf = local.new_zeros(per, S, C)
        buf[:local.shape[0]] = local
    out = local.new_empty(world * per, S, C)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    else:
        chunks = list(out.view(world, per, S, C).unbind(0))
        dist.all_gather(chunks, buf.contiguous(), group=group)
    if world * per == num_cams:
        return out
    keep = torch.cat([torch.arange(r * per, r * per + (b - a)) for r, (a, b) in enumerate(plan)]).to(out.device)
    return out.index_select(0, keep)
