TAG=${1:-r3h}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "linear or agg or softmax" ) > gpurun_out/${TAG}_pytest_a.log 2>&1; echo "pytest a exit $?"; tail -15 gpurun_out/${TAG}_pytest_a.log
( timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=line -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_b.log 2>&1; echo "pytest b exit $?"; tail -15 gpurun_out/${TAG}_pytest_b.log
