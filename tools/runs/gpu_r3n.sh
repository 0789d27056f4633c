TAG=${1:-r3n}
mkdir -p gpurun_out
for R in 0 24576 32768 0 24576; do
timeout 400 python bench.py --no-cpu-baseline --no-adaptive --conv-smem-reserve $R > gpurun_out/${TAG}_bench_res$R.json 2> gpurun_out/${TAG}_bench_res$R.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_res$R.json').read().strip().splitlines()[-1])
    print('reserve $R: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'u8', round(d['e2e_uint8']['value'],2), 'pts_head', round(d['sections_ms']['pts_head'],3), 'image', round(d['sections_ms']['image_branch_graph'],3), 'conv frac', round(d['roofline']['frac'],4))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_res$R.err').read()[-2000:])
PY
done
