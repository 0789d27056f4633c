#!/bin/bash
for c in "--shape s2 --halo 0" "--shape s2 --halo -1" "--shape s2 --halo 0 --precision bf16" "--shape s4 --halo 0" "--shape c2" "--shape fpn" "--shape s2 --halo 0 --stages 2" "--shape s2 --halo 0 --stages 4"; do python tools/conv_timeline.py $c; done
