TAG=${1:-r2x}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py -q --tb=short -p no:cacheprovider -x -k "agg" ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/${TAG}_pytest.log
{ for NQ in 1047 900; do
for V in "" "--static-grid" "--separate-weights" "--separate-weights --static-grid" "--u8" "--u8 --static-grid"; do
echo "== nq $NQ $V"; timeout 100 python tools/prof_kernels.py agg --iters 30 --nq $NQ $V | grep feat=
done; done; } > gpurun_out/${TAG}_agg_variants.txt 2>&1
cat gpurun_out/${TAG}_agg_variants.txt
