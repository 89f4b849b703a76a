#!/bin/bash
mkdir -p gpurun_out
./tools/pipe_rate > gpurun_out/r02l_pipe_rate.log 2>&1
timeout 300 python tools/prof_small.py q612 22 1 > gpurun_out/r02l_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q612 -c 1 -f -o gpurun_out/r02_q612_f32_v1 python tools/prof_small.py q612 22 1 > gpurun_out/r02l_ncu.log 2>&1
cat gpurun_out/r02l_pipe_rate.log; tail -3 gpurun_out/r02l_ncu.log
