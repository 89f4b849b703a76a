#!/bin/bash
mkdir -p gpurun_out
python tools/prof_small.py tiny10 21 1 > gpurun_out/r02i_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tiny -c 1 -f -o gpurun_out/r02_tiny10_v4 python tools/prof_small.py tiny10 21 1 > gpurun_out/r02i_ncu.log 2>&1
cat gpurun_out/r02i_plain.log; tail -n 3 gpurun_out/r02i_ncu.log
