#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "exit $?" >> gpurun_out/r02g_bench.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1
tail -n 5 gpurun_out/r02g_bench.err gpurun_out/r02g_smoke.log
