TAG=${1:-r2y}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py -q --tb=short -p no:cacheprovider -x -k "agg" ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/${TAG}_pytest.log
{ for NQ in 1047; do
for V in "" "--no-prefetch" "--static-grid" "--static-grid --no-prefetch" "--u8 --static-grid" "--u8 --static-grid --no-prefetch" "--u8"; do
echo "== nq $NQ $V"; timeout 100 python tools/prof_kernels.py agg --iters 30 --nq $NQ $V | grep feat=
done; done; } > gpurun_out/${TAG}_agg_variants.txt 2>&1
cat gpurun_out/${TAG}_agg_variants.txt
for T in 1 17 3 19 11 27; do
timeout 400 python bench.py --no-cpu-baseline --no-adaptive --agg-tune $T > gpurun_out/${TAG}_bench_agg$T.json 2> gpurun_out/${TAG}_bench_agg$T.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_agg$T.json').read().strip().splitlines()[-1])
    print('agg tune $T: value', round(d['value'],2), 'pts_head', round(d['sections_ms']['pts_head'],3), 'agg', round(d['roofline_deform_agg']['frac'],4), round(d['roofline_deform_agg']['kernel_us_per_launch'],2), 'conv', round(d['roofline']['frac'],4))
except Exception as e:
    print('bench parse failed', e)
PY
done
