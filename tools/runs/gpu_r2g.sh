TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 300 python tools/conv_frame_breakdown.py --precision fp16mx > gpurun_out/${TAG}_conv_breakdown_fp16mx.txt 2>&1; cat gpurun_out/${TAG}_conv_breakdown_fp16mx.txt
timeout 300 python tools/conv_frame_breakdown.py --precision fp16x3 > gpurun_out/${TAG}_conv_breakdown_fp16x3.txt 2>&1; head -12 gpurun_out/${TAG}_conv_breakdown_fp16x3.txt
timeout 120 python tools/prof_kernels.py conv --shape s5 --precision fp16mx --iters 10
timeout 600 python tests/tools/diag_cfg2_parity.py fp16mx > gpurun_out/${TAG}_diag_cfg2.txt 2>&1; tail -22 gpurun_out/${TAG}_diag_cfg2.txt
