TAG=${1:-r4h}
mkdir -p gpurun_out
timeout 200 python - > gpurun_out/${TAG}_resize_timing.txt 2>&1 <<'PY'
import numpy as np, torch
from far3d_b200 import imgproc, ops, _lib
dev = torch.device('cuda:0')
conf = dict(resize_lim=(0.47, 0.55), final_dim=(640, 960), final_dim_f=(640, 720), bot_pct_lim=(0.0, 0.0), rot_lim=(0.0, 0.0), rand_flip=False)
T = imgproc.AV2ResizeCropFlipRotImageV2(data_aug_conf=conf)
rng = np.random.default_rng(0)
shapes = [(2048, 1550)] + [(1550, 2048)] * 6                       # AV2 rig: portrait front centre + six ring cameras
devv = [torch.from_numpy(rng.integers(0, 256, size=hw + (3,), dtype=np.uint8)).to(dev) for hw in shapes]
K = [np.eye(4) for _ in shapes]
imgproc.prefetch_tables((1550, 2048), conf['resize_lim'], dev)
batch = torch.empty(7, 640, 960, 3, device=dev, dtype=torch.uint8)
np.random.seed(1)
plans = [T.plan([tuple(v.shape) for v in devv], [k.copy() for k in K], K)[0] for _ in range(12)]
for p in plans[:2]:
    T.apply(devv, p, out=batch)
torch.cuda.synchronize()
# GPU time of the kernels alone: the host queued ahead behind a device-side spin
torch.cuda._sleep(int(30e-3 * 1.9e9))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for p in plans[2:]:
    T.apply(devv, p, out=batch)
e1.record(); torch.cuda.synchronize()
print(f'resize + crop kernels, 7 views (random resize per view), device time with the host queued ahead: {e0.elapsed_time(e1) / 10:.3f} ms per frame')
import time
t0 = time.perf_counter()
for p in plans[2:]:
    T.apply(devv, p, out=batch)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f'host enqueue time: {(t1 - t0) * 100:.3f} ms per frame (16 launches)')
PY
cat gpurun_out/${TAG}_resize_timing.txt
( time timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'u8', round(d['e2e_uint8']['value'],3), 'lat', round(d['latency_ms_unpipelined'],3), 'conv', (round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_frame'],3)), 'agg', (round(d['roofline_deform_agg']['frac'],3), round(d['roofline_deform_agg']['kernel_us_per_launch'],1)), 'clocks', d.get('clocks'))
print('raw', d.get('e2e_raw_cameras'))
print('adaptive', d['streaming_adaptive']['value'], d['streaming_adaptive'].get('clocks'))
PY
