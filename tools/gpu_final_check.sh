# final GPU-box pass of a round, in order of importance, every step under its own timeout
# usage: bash tools/gpu_final_check.sh <tag>
TAG=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/${TAG}_pytest.log
( time timeout 120 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
head -c 5000 gpurun_out/${TAG}_bench.json; tail -4 gpurun_out/${TAG}_bench.err
( time timeout 120 python __graft_entry__.py smoke ) > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
# DRAM traffic of every conv launch of ONE frame (the third eager warm-up frame): dram bytes + duration per launch
timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_persistent --launch-skip 264 --launch-count 132 --csv --log-file gpurun_out/${TAG}_conv_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --eager > gpurun_out/${TAG}_ncu_traffic.log 2>&1; echo "ncu conv traffic exit $?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 1600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --eager > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu list exit $?"
