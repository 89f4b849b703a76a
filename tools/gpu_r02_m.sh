#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02m_*
timeout 900 python -m pytest tests/test_gpu_q612.py tests/test_gpu_sdr.py tests/test_gpu_metrics.py -x -q -m gpu > gpurun_out/r02m_pytest_q612.log 2>&1
echo "exit $?" >> gpurun_out/r02m_pytest_q612.log
timeout 300 python tools/prof_small.py q612 22 5 >> gpurun_out/r02m_time_q612.log 2>&1
timeout 300 python tools/prof_small.py q612f10 21 5 >> gpurun_out/r02m_time_q612.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q612 -c 1 -f -o gpurun_out/r02_q612_v2 python tools/prof_small.py q612 22 1 > gpurun_out/r02m_ncu.log 2>&1
tail -n 4 gpurun_out/r02m_pytest_q612.log; cat gpurun_out/r02m_time_q612.log
