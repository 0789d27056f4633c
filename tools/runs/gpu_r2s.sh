TAG=${1:-r2s}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_mx.py -q --tb=short -p no:cacheprovider -x -k "conv or mx" ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/${TAG}_pytest.log
{ for S in s4 s5 c4; do timeout 60 python tools/conv_timeline.py --shape $S --precision fp16mx; done; } > gpurun_out/${TAG}_conv_timeline.txt 2>&1
grep -v "^$" gpurun_out/${TAG}_conv_timeline.txt | tail -40
timeout 300 python tools/conv_frame_breakdown.py --precision fp16mx > gpurun_out/${TAG}_conv_frame_breakdown.txt 2>&1; tail -45 gpurun_out/${TAG}_conv_frame_breakdown.txt
timeout 400 python bench.py --no-cpu-baseline --no-adaptive > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'sections', d.get('sections_ms'), 'conv frac', d['roofline'] and round(d['roofline']['frac'],4), 'latency', d.get('latency_ms_unpipelined'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-3000:])
PY
