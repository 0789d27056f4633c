TAG=${1:-r4i}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 200 python __graft_entry__.py smoke ) > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/${TAG}_smoke.log
