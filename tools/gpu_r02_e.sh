#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_modes.py f16x3,bf16 10 > gpurun_out/r02e_time_modes.log 2>&1
timeout 600 python tools/diag_error.py 1024 f16x3 > gpurun_out/r02e_diag_error.log 2>&1
timeout 1800 python -m pytest tests/test_gpu_vt.py -x -q -m gpu > gpurun_out/r02e_pytest_vt.log 2>&1
echo "exit $?" >> gpurun_out/r02e_pytest_vt.log
cat gpurun_out/r02e_time_modes.log gpurun_out/r02e_diag_error.log; tail -n 8 gpurun_out/r02e_pytest_vt.log
