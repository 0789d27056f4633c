#!/bin/bash
python tools/prof_kernels.py conv --precision fp16x3 --iters 3 2>&1 | grep conv
python tools/prof_kernels.py conv --precision fp16 --iters 3 2>&1 | grep conv
python tools/conv_timeline.py --shape s2 --halo 0
python tools/conv_timeline.py --shape c2
