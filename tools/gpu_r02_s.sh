#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02s_*
timeout 1200 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
echo "exit $?" >> gpurun_out/r02s_bench.err
tail -n 5 gpurun_out/r02s_bench.err
