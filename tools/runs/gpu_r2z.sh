TAG=${1:-r2z}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q --tb=short -p no:cacheprovider -x -k "agg or detector or graph or decoder or layer" ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/${TAG}_pytest.log
{ for NQ in 1047 900; do
for V in "" "--u4" "--prepared" "--prepared --u4" "--prepared --work-queue"; do
echo "== nq $NQ $V"; timeout 100 python tools/prof_kernels.py agg --iters 30 --nq $NQ $V | grep "feat=\|dfa_prepare"
done; done; } > gpurun_out/${TAG}_agg_variants.txt 2>&1
cat gpurun_out/${TAG}_agg_variants.txt
timeout 400 python bench.py --no-cpu-baseline --no-adaptive > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'sections', d['sections_ms'], 'agg', round(d['roofline_deform_agg']['frac'],4), round(d['roofline_deform_agg']['kernel_us_per_launch'],2), 'conv', round(d['roofline']['frac'],4), 'latency', d['latency_ms_unpipelined'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-2000:])
PY
