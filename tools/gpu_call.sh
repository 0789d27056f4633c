TAG=r1j
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_2gpu_streams.json 2> gpurun_out/${TAG}_bench_2gpu_streams.err; echo "bench streams exit $?"
head -c 1500 gpurun_out/${TAG}_bench_2gpu_streams.json; echo; tail -3 gpurun_out/${TAG}_bench_2gpu_streams.err
