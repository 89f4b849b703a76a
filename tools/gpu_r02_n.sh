#!/bin/bash
# full GPU suite + smoke + default bench + reference arm on one B200
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02n_pytest_all.log 2>&1
echo "exit $?" >> gpurun_out/r02n_pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1
echo "exit $?" >> gpurun_out/r02n_smoke.log
timeout 1200 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
echo "exit $?" >> gpurun_out/r02n_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err
echo "exit $?" >> gpurun_out/r02n_bench_ref.err
timeout 300 python tools/time_tiny.py 21 > gpurun_out/r02n_time_tiny.log 2>&1
tail -n 4 gpurun_out/r02n_pytest_all.log gpurun_out/r02n_smoke.log gpurun_out/r02n_bench.err gpurun_out/r02n_bench_ref.err gpurun_out/r02n_time_tiny.log
