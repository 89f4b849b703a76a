#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02p_*
timeout 900 python -m pytest tests/test_gpu_tiny.py tests/test_gpu_metrics.py tests/test_gpu_vt.py -x -q -m gpu -k "tiny or metrics or Tiny" > gpurun_out/r02p_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r02p_pytest.log
timeout 300 python tools/time_tiny.py 21 > gpurun_out/r02p_time_tiny.log 2>&1
tail -n 5 gpurun_out/r02p_pytest.log; cat gpurun_out/r02p_time_tiny.log
