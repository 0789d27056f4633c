#!/bin/bash
# experiment matrix for the conv kernels (run on the GPU box)
for shape in s2 s4 c2; do
  for prec in bf16x3 bf16; do
    for cfg in "--cm 1 --halo -1 --stages 2" "--cm 1 --halo -1 --stages 3" "--cm 1 --halo -1" "--cm 2 --halo -1" "--cm 1 --halo 0" "--cm 2 --halo 0" "--cm 4 --halo 0"; do
      echo -n "[$cfg] "; python tools/prof_kernels.py conv --shape $shape --precision $prec --iters 3 $cfg 2>&1 | tail -1
    done
  done
done
