#!/bin/bash
python tools/conv_timeline.py --shape s2 --halo 0
python tools/conv_timeline.py --shape c2
python tools/conv_timeline.py --shape fpn
