TAG=${1:-r4y}
mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'u8', round(d['e2e_uint8']['value'],3), 'lat', round(d['latency_ms_unpipelined'],3), 'conv', (round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_frame'],3)), 'agg', (round(d['roofline_deform_agg']['frac'],3), round(d['roofline_deform_agg']['kernel_us_per_launch'],1)), 'clocks', d.get('clocks'))
print('adaptive', d.get('streaming_adaptive'))
print('sections', d['sections_ms'])
PY
