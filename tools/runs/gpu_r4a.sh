# round 3 (session 2): 8 epilogue warps with 32-column rounds vs the 4-warp build (libfar3d_sm100_epi4.so), same tree
TAG=${1:-r4a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/${TAG}_pytest.log
for L in epi8 epi4; do
  if [ $L = epi4 ]; then export FAR3D_LIB_PATH=$PWD/far3d_b200/lib/libfar3d_sm100_epi4.so; else unset FAR3D_LIB_PATH; fi
  timeout 200 python tools/prof_kernels.py conv --shape all --precision fp16mx --iters 10 > gpurun_out/${TAG}_conv_classes_$L.txt 2>&1
  cat gpurun_out/${TAG}_conv_classes_$L.txt
  timeout 300 python bench.py --no-cpu-baseline --no-adaptive > gpurun_out/${TAG}_bench_$L.json 2> gpurun_out/${TAG}_bench_$L.err; echo "bench $L exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$L.json').read().strip().splitlines()[-1])
    print('$L: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'lat', round(d['latency_ms_unpipelined'],3), 'sections', {k: round(v,3) for k,v in d['sections_ms'].items()}, 'conv frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_frame'],3))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_$L.err').read()[-2000:])
PY
done
