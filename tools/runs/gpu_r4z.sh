# end-of-round evidence on one B200 (final tree): tests, smoke, bench + reference arm, ncu launch list, conv DRAM traffic of one frame,
# ncu --set full of the roofline kernels, layer-class timings
TAG=${1:-r4z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 200 python __graft_entry__.py smoke ) > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench ref exit $?"
python - <<PY
import json
for f in ('${TAG}_bench.json','${TAG}_bench_reference.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'conv', d.get('roofline') and (round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_frame'],3)), 'agg', d.get('roofline_deform_agg') and (round(d['roofline_deform_agg']['frac'],3), round(d['roofline_deform_agg']['kernel_us_per_launch'],1)), 'clocks', d.get('clocks'), 'adaptive', d.get('streaming_adaptive') and d['streaming_adaptive'].get('value'))
    except Exception as e:
        print(f, 'parse failed', e)
PY
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-count 12000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-adaptive --eager > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list exit $?"
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv --title "round 2 end (second session), fp16mx, one eager cfg-2 frame (ncu launch list of: python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-adaptive --eager)" > gpurun_out/${TAG}_launches_summary.txt; head -12 gpurun_out/${TAG}_launches_summary.txt
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_persistent --csv --log-file gpurun_out/${TAG}_conv_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-adaptive --eager > gpurun_out/${TAG}_ncu_traffic.log 2>&1; echo "ncu conv traffic exit $?"
python tools/conv_traffic.py gpurun_out/${TAG}_conv_traffic.csv --mode 2 > gpurun_out/${TAG}_conv_traffic_one_frame.txt; head -3 gpurun_out/${TAG}_conv_traffic_one_frame.txt
for S in s2 s4 c4; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_persistent -s 1 -c 1 -f -o /tmp/${TAG}_conv_${S} python tools/prof_kernels.py conv --shape $S --precision fp16mx --iters 1 > gpurun_out/${TAG}_ncu_conv_${S}.log 2>&1
  ncu -i /tmp/${TAG}_conv_${S}.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_${S}_fp16mx_raw.csv 2>/dev/null; echo "ncu conv $S exit $?"
done
python tools/ncu_summary.py gpurun_out/${TAG}_conv_s2_fp16mx_raw.csv gpurun_out/${TAG}_conv_s4_fp16mx_raw.csv gpurun_out/${TAG}_conv_c4_fp16mx_raw.csv > gpurun_out/${TAG}_conv_ncu_summary.txt 2>&1
timeout 200 python tools/prof_kernels.py conv --shape all --precision fp16mx --iters 10 > gpurun_out/${TAG}_conv_layer_classes.txt 2>&1
timeout 200 python tools/conv_frame_breakdown.py --precision fp16mx > gpurun_out/${TAG}_conv_frame_breakdown_fp16mx.txt 2>&1
head -30 gpurun_out/${TAG}_conv_frame_breakdown_fp16mx.txt
