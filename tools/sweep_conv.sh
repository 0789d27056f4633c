#!/bin/bash
python tools/prof_kernels.py conv --precision bf16x3 --iters 3 2>&1 | grep conv
