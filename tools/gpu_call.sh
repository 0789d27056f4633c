TAG=r1i
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_ref_golden.py -m gpu -q --tb=short -p no:cacheprovider -x -k "deform or aggregation or cfg2" ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/${TAG}_pytest.log
for V in "4 " "8 " "4 --narrow" "8 --narrow" "4 --fp16"; do set -- $V; timeout 120 python tools/prof_kernels.py agg --warps $1 $2 --iters 20 | grep -v warm; done > gpurun_out/${TAG}_agg_variants.txt 2>&1
cat gpurun_out/${TAG}_agg_variants.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deform_agg --launch-skip 2 -c 1 -f -o gpurun_out/${TAG}_agg python tools/prof_kernels.py agg > gpurun_out/${TAG}_agg.log 2>&1; echo "ncu agg exit $?"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'sections', d.get('sections_ms'), 'conv frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_frame'],3), 'agg us', round(d['roofline_deform_agg']['kernel_us_per_launch'],1), 'agg frac', round(d['roofline_deform_agg']['frac'],3))
PY
