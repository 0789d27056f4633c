"""Camera-sharded frame (the north_star's all-gather design) on the GPU: two ranks run the image
branch on their camera slices, exchange the flattened feature maps and the dense 2D-head maps, and decode replicated.
With >= 2 GPUs the ranks use NCCL on separate devices; on a one-GPU box both ranks share cuda:0 and exchange through gloo
(host staging) - the product path on the 8-GPU box is the NCCL one (bench.py --shard cameras)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, num_cams, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    try:
        import torch.distributed as dist
        from far3d_b200 import synthetic
        from far3d_b200.parallel import CameraShardedFar3D
        from helpers import build_oracle, build_product, model_cfg, rel_err, to_dev
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        ndev = torch.cuda.device_count()
        dev = torch.device('cuda', rank % ndev)
        torch.cuda.set_device(dev)
        backend = 'nccl' if ndev >= world else 'gloo'
        dist.init_process_group(backend, rank=rank, world_size=world)
        mc = model_cfg(num_cams=num_cams)
        o = build_oracle(mc, seed=1)
        p = build_product(mc, o.state_dict(), dev)
        frames = [synthetic.make_frame((num_cams, 128, 192), f) for f in range(2)]
        keys = ('all_cls_scores', 'all_bbox_preds', 'feat_flatten', 'reference_points2d')
        single = []
        for metas, data in frames:                                   # the whole rig on one GPU
            res = p.simple_test(metas, **to_dev(data, dev))
            single.append(({k: p.last_outs[k].clone() for k in keys}, res[0]['pts_bbox']['scores_3d'].clone()))
        mem_single = p.pts_bbox_head.memory_embedding.clone()
        p.prev_scene_token = None
        sh = CameraShardedFar3D(p)
        errs, exact = [], True
        for (metas, data), (want, scores) in zip(frames, single):
            res = sh.simple_test(metas, **to_dev(data, dev))
            for k in keys:
                got = p.last_outs[k]
                assert got.shape == want[k].shape, (k, got.shape, want[k].shape)
                errs.append(rel_err(got, want[k]))
                exact &= bool(torch.equal(got, want[k]))
            errs.append(rel_err(res[0]['pts_bbox']['scores_3d'], scores))
        errs.append(rel_err(p.pts_bbox_head.memory_embedding, mem_single))
        # every rank ends with the same result
        mine = p.last_outs['all_cls_scores'].float().cpu()
        both = [torch.empty_like(mine) for _ in range(world)]
        if backend == 'nccl':
            g = [torch.empty_like(p.last_outs['all_cls_scores']) for _ in range(world)]
            dist.all_gather(g, p.last_outs['all_cls_scores'].contiguous())
            both = [t.cpu() for t in g]
        else:
            dist.all_gather(both, mine)
        same = all(torch.equal(both[0], t) for t in both)
        # host entry: only the local camera slice is uploaded
        p.prev_scene_token = None
        out, h2d, d2h = sh.infer(frames[0][0], **frames[0][1])
        a, b = sh.camera_range(num_cams)
        assert h2d < (b - a) * 3 * 128 * 192 * 4 + 4096 and d2h > 0
        errs.append(rel_err(out[0]['pts_bbox']['scores_3d'], single[0][1]))
        q.put((rank, max(errs), exact, same, backend, sh.last_gather_bytes))
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc(), False, False, '?', 0))


@pytest.mark.parametrize('num_cams', [2, 3])
def test_camera_sharded_frame_equals_single_gpu(cuda, lib_built, num_cams):
    """2 ranks; 2 cameras (1 + 1) and 3 cameras (ragged 2 + 1)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (os.getpid() + num_cams) % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, num_cams, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=600) for _ in range(2))
    [p.join(60) for p in ps]
    for rank, err, exact, same, backend, nbytes in res:
        assert not isinstance(err, str), err
        assert err < 1e-5, (rank, err)
        assert same, 'ranks disagree'
        assert nbytes > 0
    print('camera-sharded == single GPU:', [(r[1], 'bit-exact' if r[2] else 'within tol', r[4]) for r in res])
