TAG=${1:-r2o}
mkdir -p gpurun_out
for MK in 1 0; do
FAR3D_MEMORY_KERNELS=$MK timeout 500 python bench.py --no-cpu-baseline --no-adaptive > gpurun_out/${TAG}_bench_mk$MK.json 2> gpurun_out/${TAG}_bench_mk$MK.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_mk$MK.json').read().strip().splitlines()[-1])
print('memory kernels $MK: value', round(d['value'],2), 'sections', d.get('sections_ms'), 'eager', d.get('sections_eager_ms'), 'latency', d.get('latency_ms_unpipelined'))
PY
done
