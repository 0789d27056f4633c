TAG=${1:-r3b}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/${TAG}_pytest_gpu.log
{ for S in c3 c5 c2 c4; do for BN in 0 128 192 224; do timeout 100 python tools/prof_kernels.py conv --shape $S --precision fp16mx --iters 10 --bn $BN | sed "s/^/bn $BN: /"; done; done; } > gpurun_out/${TAG}_conv_bn.txt 2>&1
cat gpurun_out/${TAG}_conv_bn.txt
