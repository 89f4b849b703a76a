#!/bin/bash
mkdir -p gpurun_out
for mode in f16x3 tf32x3 bf16; do
  timeout 300 python tools/diag_position.py $mode 4096 4 > gpurun_out/r02b_diag_${mode}_4096.log 2>&1
  timeout 300 python tools/diag_position.py $mode 4097 3 > gpurun_out/r02b_diag_${mode}_4097.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_vt.py -q -m gpu -k "repeated_passes or full_batch or raw_u8 or pageable or returns_before or reserve or range_fallback or rejected" > gpurun_out/r02b_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r02b_pytest.log
cat gpurun_out/r02b_diag_*.log; tail -n 30 gpurun_out/r02b_pytest.log
