#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02t_*
for c in 8192 16384 32768 65536; do MDC_HOST_CHUNK=$c timeout 300 python tools/e2e_small.py >> gpurun_out/r02t_e2e.log 2>&1; done
cat gpurun_out/r02t_e2e.log
