#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tiny.py tests/test_gpu_metrics.py -x -q -m gpu > gpurun_out/r02h_pytest_tiny.log 2>&1
echo "exit $?" >> gpurun_out/r02h_pytest_tiny.log
timeout 300 python tools/time_tiny.py 21 > gpurun_out/r02h_time_tiny.log 2>&1
MDC_TINY_VARIANT=3 timeout 300 python tools/time_tiny.py 21 >> gpurun_out/r02h_time_tiny.log 2>&1
tail -n 6 gpurun_out/r02h_pytest_tiny.log; cat gpurun_out/r02h_time_tiny.log
