#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02v_*
timeout 1500 python -m pytest tests/test_gpu_vt.py -x -q -m gpu > gpurun_out/r02v_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r02v_pytest.log
timeout 300 python tools/time_modes.py f16x3 20 > gpurun_out/r02v_time_modes.log 2>&1
MDC_VT_NO_SPLIT=1 timeout 300 python tools/time_modes.py f16x3 20 >> gpurun_out/r02v_time_modes.log 2>&1
timeout 300 python tools/time_modes.py f16x3 20 >> gpurun_out/r02v_time_modes.log 2>&1
MDC_VT_NO_SPLIT=1 timeout 300 python tools/time_modes.py f16x3 20 >> gpurun_out/r02v_time_modes.log 2>&1
tail -n 5 gpurun_out/r02v_pytest.log; cat gpurun_out/r02v_time_modes.log
