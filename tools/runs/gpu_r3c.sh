TAG=${1:-r3c}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'u8', round(d['e2e_uint8']['value'],2), 'adaptive', d['streaming_adaptive'] and round(d['streaming_adaptive']['value'],2), 'sections', d['sections_ms'], 'eager', d['sections_eager_ms'], 'agg', round(d['roofline_deform_agg']['frac'],4), round(d['roofline_deform_agg']['kernel_us_per_launch'],2), 'conv', round(d['roofline']['frac'],4), 'latency', d['latency_ms_unpipelined'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-2000:])
PY
