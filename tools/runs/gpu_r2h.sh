TAG=${1:-r2h}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_mx.py tests/test_gpu_ops.py -q --tb=short -p no:cacheprovider -x -k "conv or mx or linear" ) > gpurun_out/${TAG}_pytest_conv.log 2>&1; echo "pytest conv exit $?"
tail -4 gpurun_out/${TAG}_pytest_conv.log
( timeout 300 python tools/prof_kernels.py conv --shape all --precision fp16mx --iters 10 ) > gpurun_out/${TAG}_conv_classes.txt 2>&1
cat gpurun_out/${TAG}_conv_classes.txt
for P in fp16mx; do
  timeout 400 python bench.py --no-cpu-baseline --precision $P > gpurun_out/${TAG}_bench_$P.json 2> gpurun_out/${TAG}_bench_$P.err; echo "bench $P exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$P.json').read().strip().splitlines()[-1])
    print('$P', 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'sections', d.get('sections_ms'), 'conv frac', d['roofline'] and round(d['roofline']['frac'],4), 'conv ms', d['roofline'] and round(d['roofline']['kernel_ms_per_frame'],3), 'agg us', d['roofline_deform_agg'] and round(d['roofline_deform_agg']['kernel_us_per_launch'],1))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_$P.err').read()[-3000:])
PY
done
