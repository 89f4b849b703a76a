#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_error.py 1024 f16x3 > gpurun_out/r02d_diag_error.log 2>&1
timeout 300 python tools/time_modes.py f16x3,bf16 10 > gpurun_out/r02d_time_modes.log 2>&1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02d_pytest_all.log 2>&1
echo "exit $?" >> gpurun_out/r02d_pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
echo "exit $?" >> gpurun_out/r02d_bench.err
cat gpurun_out/r02d_diag_error.log gpurun_out/r02d_time_modes.log; tail -n 25 gpurun_out/r02d_pytest_all.log; tail -n 5 gpurun_out/r02d_bench.err; head -c 3000 gpurun_out/r02d_bench.json
