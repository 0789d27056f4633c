TAG=${1:-r2q}
mkdir -p gpurun_out
{ for S in s4 s5 s2 c4 s3 fpn; do timeout 60 python tools/conv_timeline.py --shape $S --precision fp16mx; done; timeout 60 python tools/conv_timeline.py --shape s4 --precision fp16x3; } > gpurun_out/${TAG}_conv_timeline.txt 2>&1
cat gpurun_out/${TAG}_conv_timeline.txt
timeout 120 python tools/prof_kernels.py conv --shape c5 --precision fp16mx --iters 10
