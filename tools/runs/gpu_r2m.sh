TAG=${1:-r2m}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_ref_golden.py tests/test_gpu_model.py -q --tb=short -p no:cacheprovider -k "decode or coder or detector or cfg2 or device_proposals or interleaved or graph" ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --no-cpu-baseline --no-adaptive > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'sections', d.get('sections_ms'), 'conv frac', d['roofline'] and round(d['roofline']['frac'],4), 'latency', d.get('latency_ms_unpipelined'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-3000:])
PY
