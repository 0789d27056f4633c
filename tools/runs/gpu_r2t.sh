TAG=${1:-r2t}
mkdir -p gpurun_out
FAR3D_NVCC_EXTRA=-DFAR3D_CONV_WAITSTATS python -c "
from far3d_b200 import build; build.build(force=True)" > /dev/null 2>&1; echo "rebuild exit $?"
{ for S in s2 s3 s4 s4b c4 c3 c2 fpn stem2 s5; do timeout 60 python tools/conv_timeline.py --shape $S --precision fp16mx --ghz 1.92; done; } > gpurun_out/${TAG}_conv_waits.txt 2>&1
cat gpurun_out/${TAG}_conv_waits.txt
