#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_q612.py -x -q -m gpu > gpurun_out/r02k_pytest_q612.log 2>&1
echo "exit $?" >> gpurun_out/r02k_pytest_q612.log
for v in 0 1; do
  echo "MDC_Q612_VARIANT=$v" >> gpurun_out/r02k_time_q612.log
  MDC_Q612_VARIANT=$v timeout 300 python tools/prof_small.py q612 22 5 >> gpurun_out/r02k_time_q612.log 2>&1
  MDC_Q612_VARIANT=$v timeout 300 python tools/prof_small.py q612f10 21 5 >> gpurun_out/r02k_time_q612.log 2>&1
done
tail -n 8 gpurun_out/r02k_pytest_q612.log; cat gpurun_out/r02k_time_q612.log
