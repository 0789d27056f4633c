TAG=${1:-r4g}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_ref_golden.py -m gpu -q -k "raw_camera or uint8 or resize" --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/${TAG}_pytest.log
( time timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'u8', round(d['e2e_uint8']['value'],3), 'lat', round(d['latency_ms_unpipelined'],3), 'conv', (round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_frame'],3)), 'agg', (round(d['roofline_deform_agg']['frac'],3), round(d['roofline_deform_agg']['kernel_us_per_launch'],1)), 'clocks', d.get('clocks'))
print('raw', d.get('e2e_raw_cameras'))
print('adaptive', d['streaming_adaptive']['value'], d['streaming_adaptive'].get('clocks'))
PY
tail -3 gpurun_out/${TAG}_bench.err
