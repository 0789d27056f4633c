# resident weights for single-chunk halo layers; weight exponent window of the fp16mx format (row-error distributions at cfg-2)
TAG=${1:-r4d}
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_mx.py tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_conv.log 2>&1
rc=$?; echo "conv pytest exit $rc"; tail -6 gpurun_out/${TAG}_pytest_conv.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 200 python tools/prof_kernels.py conv --shape all --precision fp16mx --iters 10 > gpurun_out/${TAG}_conv_classes.txt 2>&1
cat gpurun_out/${TAG}_conv_classes.txt
for T in 5 3 7; do
  echo "== FAR3D_MX_W_TOP=$T"
  FAR3D_MX_W_TOP=$T timeout 300 python tests/tools/diag_cfg2_parity.py fp16mx > gpurun_out/${TAG}_diag_wtop$T.txt 2>&1
  cat gpurun_out/${TAG}_diag_wtop$T.txt | tail -22
done
( time timeout 600 python -m pytest tests/test_gpu_ref_golden.py -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_golden.log 2>&1; echo "golden pytest exit $?"
tail -8 gpurun_out/${TAG}_pytest_golden.log
