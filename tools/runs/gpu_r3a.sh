TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-count 12000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-adaptive --eager > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list exit $?"
wc -l gpurun_out/${TAG}_launches.csv
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv --title "round 2, fp16mx, cfg-2 eager frame (ncu launch list)" > gpurun_out/${TAG}_launches_summary.txt; cat gpurun_out/${TAG}_launches_summary.txt
