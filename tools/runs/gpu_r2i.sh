TAG=${1:-r2i}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_gpu.log | tail -5
grep -E "^(FAILED|ERROR)|scale error" gpurun_out/${TAG}_pytest_gpu.log | head -40
for V in "4 " "2 " "8 "; do set -- $V; timeout 120 python tools/prof_kernels.py agg --warps $1 --iters 20; done > gpurun_out/${TAG}_agg_variants.txt 2>&1
cat gpurun_out/${TAG}_agg_variants.txt
