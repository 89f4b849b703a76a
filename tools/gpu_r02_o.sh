#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02o_*
timeout 1500 python -m pytest tests/test_gpu_vt.py tests/test_gpu_q612.py -x -q -m gpu > gpurun_out/r02o_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r02o_pytest.log
timeout 300 python tools/time_modes.py f16x3 20 > gpurun_out/r02o_time_modes.log 2>&1
timeout 300 python tools/time_modes.py f16x3 20 >> gpurun_out/r02o_time_modes.log 2>&1
timeout 300 python tools/prof_small.py q612f10 21 5 >> gpurun_out/r02o_time_modes.log 2>&1
tail -n 5 gpurun_out/r02o_pytest.log; cat gpurun_out/r02o_time_modes.log
