# N-GPU pass (gpurun --gpus N): NCCL camera-shard test, bench in both sharding modes
# usage: bash tools/gpu_multi_check.sh <tag> <N>
TAG=${1:-rX}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi_multi.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider -s ) > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -6 gpurun_out/${TAG}_pytest_multi.log
for MODE in streams cameras; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --shard $MODE --no-cpu-baseline > gpurun_out/${TAG}_bench_${N}gpu_${MODE}.json 2> gpurun_out/${TAG}_bench_${N}gpu_${MODE}.err; echo "bench $MODE exit $?"
  head -c 2600 gpurun_out/${TAG}_bench_${N}gpu_${MODE}.json; echo; tail -3 gpurun_out/${TAG}_bench_${N}gpu_${MODE}.err
done
