#!/bin/bash
for prec in bf16x3 bf16; do
  python tools/prof_kernels.py conv --precision $prec --iters 3 2>&1 | grep conv
done
python tools/conv_timeline.py --shape s2 --halo 0
python tools/conv_timeline.py --shape s2 --halo -1
python tools/conv_timeline.py --shape s2 --halo 0 --precision bf16
python tools/conv_timeline.py --shape s4 --halo 0
