TAG=${1:-r2v}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q --tb=short -p no:cacheprovider -x -k "agg or detector or graph" ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/${TAG}_pytest.log
{ for NQ in 900 1047; do
timeout 100 python tools/prof_kernels.py agg --iters 30 --nq $NQ
timeout 100 python tools/prof_kernels.py agg --iters 30 --nq $NQ --static-grid
timeout 100 python tools/prof_kernels.py agg --iters 30 --nq $NQ --warps 2
timeout 100 python tools/prof_kernels.py agg --iters 30 --nq $NQ --warps 8
done; } > gpurun_out/${TAG}_agg_variants.txt 2>&1
cat gpurun_out/${TAG}_agg_variants.txt
