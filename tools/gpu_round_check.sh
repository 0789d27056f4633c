mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r1c_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/r1c_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r1c_pytest.log
( time timeout 600 python bench.py ) > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; echo "bench exit $?"
cat gpurun_out/r1c_bench.json | head -c 3000
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r1c_bench_ref.json 2> gpurun_out/r1c_bench_ref.err; echo "ref exit $?"
cat gpurun_out/r1c_bench_ref.json | head -c 1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 2400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile --eager > gpurun_out/r1c_ncu.log 2>&1; echo "ncu list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_persistent --launch-skip 2 -c 1 -f -o gpurun_out/r1c_conv_s3 python tools/prof_kernels.py conv --shape s3 > gpurun_out/r1c_conv_s3.log 2>&1; echo "ncu conv exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deform_agg --launch-skip 2 -c 1 -f -o gpurun_out/r1c_agg python tools/prof_kernels.py agg > gpurun_out/r1c_agg.log 2>&1; echo "ncu agg exit $?"
for s in s2 s3 s4 s4b c3 c4 s5 c5 fpn; do timeout 120 python tools/prof_kernels.py conv --shape $s; done > gpurun_out/r1c_conv_classes.txt 2>&1
timeout 120 python tools/prof_kernels.py agg >> gpurun_out/r1c_conv_classes.txt 2>&1
tail -15 gpurun_out/r1c_conv_classes.txt
