# one GPU-box pass: tests, bench (both arms), parity distribution at full size, ncu launch list + full captures
# usage: bash tools/gpu_round_check.sh <tag> [quick]
TAG=${1:-rX}; QUICK=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/${TAG}_pytest.log
timeout 500 python tests/tools/diag_cfg2_parity.py > gpurun_out/${TAG}_diag_cfg2.txt 2>&1; grep -v "^$" gpurun_out/${TAG}_diag_cfg2.txt | tail -40
( time timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
head -c 4000 gpurun_out/${TAG}_bench.json
if [ -z "$QUICK" ]; then
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref exit $?"
timeout 600 python bench.py --precision fp16 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp16.json 2> gpurun_out/${TAG}_bench_fp16.err; head -c 2500 gpurun_out/${TAG}_bench_fp16.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 2400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile --eager > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_persistent --launch-skip 2 -c 1 -f -o gpurun_out/${TAG}_conv_s3 python tools/prof_kernels.py conv --shape s3 > gpurun_out/${TAG}_conv_s3.log 2>&1; echo "ncu conv exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deform_agg --launch-skip 2 -c 1 -f -o gpurun_out/${TAG}_agg python tools/prof_kernels.py agg > gpurun_out/${TAG}_agg.log 2>&1; echo "ncu agg exit $?"
for s in s2 s3 s4 s4b c3 c4 s5 c5 fpn; do timeout 120 python tools/prof_kernels.py conv --shape $s; done > gpurun_out/${TAG}_conv_classes.txt 2>&1
timeout 120 python tools/prof_kernels.py agg >> gpurun_out/${TAG}_conv_classes.txt 2>&1
tail -12 gpurun_out/${TAG}_conv_classes.txt
fi
