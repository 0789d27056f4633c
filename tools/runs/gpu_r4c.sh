# fp16mx on plain kind::f8f6f4 (pre-scaled fp16 weight plane, no scale factors in TMEM) -> 256-column N tiles
TAG=${1:-r4c}
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_mx.py tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_conv.log 2>&1
rc=$?; echo "conv pytest exit $rc"; tail -15 gpurun_out/${TAG}_pytest_conv.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 200 python tools/prof_kernels.py conv --shape all --precision fp16mx --iters 10 > gpurun_out/${TAG}_conv_classes.txt 2>&1
cat gpurun_out/${TAG}_conv_classes.txt
( for S in c2 c3 c5 fpn fpn40 lat3 lat4; do timeout 100 python tools/prof_kernels.py conv --shape $S --precision fp16mx --iters 10 --bn 128; done
  for S in fpn40 lat4 c5; do timeout 100 python tools/prof_kernels.py conv --shape $S --precision fp16mx --iters 10 --bn 256; done
  timeout 100 python tools/prof_kernels.py conv --shape c4 --precision fp16mx --iters 10 --bn 192 ) > gpurun_out/${TAG}_conv_classes_forced_bn.txt 2>&1
echo "== forced bn (128 x7, 256 x3, c4 192)"; cat gpurun_out/${TAG}_conv_classes_forced_bn.txt
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-adaptive > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'lat', round(d['latency_ms_unpipelined'],3), 'sections', {k: round(v,3) for k,v in d['sections_ms'].items()}, 'conv frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_frame'],3))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-2000:])
PY
