# quick GPU-box pass: tests, aggregation kernel variants, bench with conv smem-reserve variants
# usage: bash tools/gpu_quick.sh <tag>
TAG=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -40 gpurun_out/${TAG}_pytest.log
for V in "4 " "8 " "4 --narrow"; do set -- $V; timeout 120 python tools/prof_kernels.py agg --warps $1 $2 --iters 20; done > gpurun_out/${TAG}_agg_variants.txt 2>&1
timeout 120 python tools/prof_kernels.py agg --fp16 --iters 20 >> gpurun_out/${TAG}_agg_variants.txt 2>&1
cat gpurun_out/${TAG}_agg_variants.txt
for R in 0 32768 49152; do
  timeout 300 python bench.py --no-cpu-baseline --conv-smem-reserve $R > gpurun_out/${TAG}_bench_reserve${R}.json 2> gpurun_out/${TAG}_bench_reserve${R}.err; echo "bench reserve $R exit $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_reserve${R}.json').read().strip().splitlines()[-1])
print('reserve', $R, 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'sections', d.get('sections_ms'), 'conv frac', d['roofline'] and round(d['roofline']['frac'],4), 'conv ms', d['roofline'] and round(d['roofline']['kernel_ms_per_frame'],3), 'agg us', d['roofline_deform_agg'] and round(d['roofline_deform_agg']['kernel_us_per_launch'],1), 'agg frac', d['roofline_deform_agg'] and round(d['roofline_deform_agg']['frac'],3))
PY
done
for s in s2 s3 s4; do timeout 120 python tools/prof_kernels.py conv --shape $s; done; timeout 120 python tools/prof_kernels.py misc > gpurun_out/${TAG}_conv_classes.txt 2>&1; tail -5 gpurun_out/${TAG}_conv_classes.txt
