TAG=${1:-r2n}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/${TAG}_pytest_gpu.log
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'u8', d['e2e_uint8'] and round(d['e2e_uint8']['value'],2), 'sections', d.get('sections_ms'), 'conv frac', d['roofline'] and round(d['roofline']['frac'],4), 'latency', d.get('latency_ms_unpipelined'), 'adaptive', d.get('streaming_adaptive') and d['streaming_adaptive'].get('value'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-3000:])
PY
