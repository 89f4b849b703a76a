#!/bin/bash
# round 2, call A: correctness of the new fp16x3 mode + first timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_vt.py -x -q -m gpu -k "f16x3 or tf32x3_mode or bf16_mode" > gpurun_out/r02a_pytest_vt_quick.log 2>&1
echo "exit $?" >> gpurun_out/r02a_pytest_vt_quick.log
timeout 300 python tools/time_modes.py f16x3,bf16,tf32x3 10 > gpurun_out/r02a_time_modes.log 2>&1
echo "exit $?" >> gpurun_out/r02a_time_modes.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02a_pytest_all.log 2>&1
echo "exit $?" >> gpurun_out/r02a_pytest_all.log
tail -5 gpurun_out/r02a_pytest_vt_quick.log gpurun_out/r02a_time_modes.log gpurun_out/r02a_pytest_all.log
