#!/bin/bash
# experiment matrix for the conv kernel (run on the GPU box)
for prec in bf16x3 bf16; do
  python tools/prof_kernels.py conv --precision $prec --iters 3 2>&1 | grep conv
  for shape in s2 s4 c2; do
    for cfg in "--halo -1" "--halo 0 --stages 2" "--halo 0 --stages 3"; do
      echo -n "[$cfg] "; python tools/prof_kernels.py conv --shape $shape --precision $prec --iters 3 $cfg 2>&1 | tail -1
    done
  done
done
python tools/conv_timeline.py --shape s2 --halo 0
python tools/conv_timeline.py --shape s2 --halo -1
python tools/conv_timeline.py --shape s2 --halo 0 --precision bf16
