# 8 epilogue warps: bit-level comparison with the 4-warp build, the failing test repeated, full suite, PDL on/off
TAG=${1:-r4b}
mkdir -p gpurun_out
E4=$PWD/far3d_b200/lib/libfar3d_sm100_epi4.so
( echo "== epi8"; timeout 200 python tools/detector_hash.py
  echo "== epi8 pdl"; timeout 200 python tools/detector_hash.py --pdl 1
  echo "== epi8 eager pdl"; timeout 200 python tools/detector_hash.py --pdl 1 --eager
  echo "== epi4"; FAR3D_LIB_PATH=$E4 timeout 200 python tools/detector_hash.py ) > gpurun_out/${TAG}_hash.txt 2>&1
cat gpurun_out/${TAG}_hash.txt | tail -30
for i in 1 2 3; do timeout 200 python -m pytest tests/test_gpu_model.py -m gpu -q -k hoisted -p no:cacheprovider 2>&1 | tail -2; done
echo "== epi4 same test"; FAR3D_LIB_PATH=$E4 timeout 200 python -m pytest tests/test_gpu_model.py -m gpu -q -k hoisted -p no:cacheprovider 2>&1 | tail -2
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/${TAG}_pytest.log
for P in 0 1; do
  timeout 300 python bench.py --no-cpu-baseline --no-adaptive --conv-pdl $P > gpurun_out/${TAG}_bench_pdl$P.json 2> gpurun_out/${TAG}_bench_pdl$P.err; echo "bench pdl $P exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_pdl$P.json').read().strip().splitlines()[-1])
    print('pdl $P: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'lat', round(d['latency_ms_unpipelined'],3), 'sections', {k: round(v,3) for k,v in d['sections_ms'].items()}, 'conv frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_frame'],3))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_pdl$P.err').read()[-2000:])
PY
done
