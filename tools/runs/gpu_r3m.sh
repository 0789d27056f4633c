TAG=${1:-r3m}
mkdir -p gpurun_out
for LM in 1 0 1 0; do
timeout 400 python bench.py --no-cpu-baseline --no-adaptive --linear-mma $LM > gpurun_out/${TAG}_bench_lm$LM.json 2> gpurun_out/${TAG}_bench_lm$LM.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_lm$LM.json').read().strip().splitlines()[-1])
    print('linear_mma $LM: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'u8', round(d['e2e_uint8']['value'],2), 'pts_head', round(d['sections_ms']['pts_head'],3), 'image', round(d['sections_ms']['image_branch_graph'],3), 'latency', round(d['latency_ms_unpipelined'],3), 'launches', d['gpu_launches'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_lm$LM.err').read()[-2000:])
PY
done
