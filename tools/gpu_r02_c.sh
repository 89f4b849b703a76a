#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_error.py 1024 > gpurun_out/r02c_diag_error.log 2>&1
timeout 300 python tools/diag_submit.py > gpurun_out/r02c_diag_submit.log 2>&1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_pytest_all.log 2>&1
echo "exit $?" >> gpurun_out/r02c_pytest_all.log
cat gpurun_out/r02c_diag_error.log gpurun_out/r02c_diag_submit.log; tail -n 25 gpurun_out/r02c_pytest_all.log
