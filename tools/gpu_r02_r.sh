#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02r_*
timeout 1500 python -m pytest tests/test_gpu_tiny.py tests/test_gpu_q612.py tests/test_gpu_sdr.py tests/test_gpu_metrics.py -x -q -m gpu > gpurun_out/r02r_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r02r_pytest.log
timeout 600 python -m pytest tests/test_gpu_vt.py -x -q -m gpu -k "raw or async or pageable" >> gpurun_out/r02r_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r02r_pytest.log
timeout 300 python tools/time_tiny.py 21 > gpurun_out/r02r_time.log 2>&1
timeout 300 python tools/prof_small.py q612 22 5 >> gpurun_out/r02r_time.log 2>&1
tail -n 12 gpurun_out/r02r_pytest.log; cat gpurun_out/r02r_time.log
