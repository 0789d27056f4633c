# resize / crop kernels (parity + timing), bench with the queued-ahead per-launch event pass
TAG=${1:-r4e}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_ref_golden.py -m gpu -q -k "resize or normalisation" --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_resize.log 2>&1; echo "resize pytest exit $?"
tail -12 gpurun_out/${TAG}_pytest_resize.log
timeout 200 python - > gpurun_out/${TAG}_resize_timing.txt 2>&1 <<'PY'
import numpy as np, torch
from far3d_b200 import imgproc, ops, _lib
dev = torch.device('cuda:0')
conf = dict(resize_lim=(0.47, 0.55), final_dim=(640, 960), final_dim_f=(640, 720), bot_pct_lim=(0.0, 0.0), rot_lim=(0.0, 0.0), rand_flip=False)
T = imgproc.AV2ResizeCropFlipRotImageV2(data_aug_conf=conf)
rng = np.random.default_rng(0)
shapes = [(2048, 1550)] + [(1550, 2048)] * 6                       # AV2 rig: portrait front centre + six ring cameras
host = [torch.from_numpy(rng.integers(0, 256, size=hw + (3,), dtype=np.uint8)).pin_memory() for hw in shapes]
devv = [h.to(dev) for h in host]
K = [np.eye(4) for _ in shapes]
def run(src):
    np.random.seed(1)
    d, _ = imgproc.frame_from_cameras(src, K, K, T, dev)
    return ops.normalize_u8(d['img'], np.float32([103.53, 116.28, 123.675]), np.float32([57.375, 57.12, 58.395]))
for name, src in (('views resident in HBM', devv), ('views in pinned host memory (66.7 MB H2D)', host)):
    for _ in range(3): run(src)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): x = run(src)
    e1.record(); torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1) / 10:.3f} ms per 7-camera frame (resize + crop + normalise -> {tuple(x.shape)}), {(_lib.launch_count() - n0) // 10} launches')
PY
cat gpurun_out/${TAG}_resize_timing.txt
for Q in 12 0; do
timeout 300 python bench.py --no-cpu-baseline --no-adaptive --profile-queue-ms $Q > gpurun_out/${TAG}_bench_q$Q.json 2> gpurun_out/${TAG}_bench_q$Q.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_q$Q.json').read().strip().splitlines()[-1])
    print('queue $Q ms: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'lat', round(d['latency_ms_unpipelined'],3), 'sections', {k: round(v,3) for k,v in d['sections_ms'].items()}, 'eager', {k: round(v,3) for k,v in d['sections_eager_ms'].items()}, 'conv frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_frame'],3), 'agg', round(d['roofline_deform_agg']['frac'],3), round(d['roofline_deform_agg']['kernel_us_per_launch'],1))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_q$Q.err').read()[-2000:])
PY
done
