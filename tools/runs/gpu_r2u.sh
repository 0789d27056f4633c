TAG=${1:-r2u}
mkdir -p gpurun_out
for P in 1 3 5; do
FAR3D_NVCC_EXTRA="-DFAR3D_CONV_WAITSTATS -DFAR3D_UM_PRODUCERS=$P" python -c "
from far3d_b200 import build; build.build(force=True)" > /dev/null 2>&1; echo "rebuild exit $?"
{ echo "===== producers $P"; for S in s2 s4b c3; do timeout 60 python tools/conv_timeline.py --shape $S --precision fp16mx --ghz 1.92; done; } >> gpurun_out/${TAG}_conv_waits_producers.txt 2>&1
done
grep -v "^$" gpurun_out/${TAG}_conv_waits_producers.txt | grep "=====\|kernel\|blocked\|busy"
