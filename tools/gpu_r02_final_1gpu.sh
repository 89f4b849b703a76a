#!/bin/bash
# round-2 final single-GPU evidence: suite, smoke, bench (both arms), launch list of the bench command, ncu captures
mkdir -p gpurun_out
rm -f gpurun_out/r02final_*
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02final_pytest_all.log 2>&1
echo "exit $?" >> gpurun_out/r02final_pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02final_smoke.log 2>&1
echo "exit $?" >> gpurun_out/r02final_smoke.log
timeout 1200 python bench.py > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err
echo "exit $?" >> gpurun_out/r02final_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02final_bench_ref.json 2> gpurun_out/r02final_bench_ref.err
echo "exit $?" >> gpurun_out/r02final_bench_ref.err
# launch list of the bench command (after it exited 0 without ncu)
timeout 900 python bench.py --steps 2 --warmup 3 --skip-other > gpurun_out/r02final_bench_short.json 2>/dev/null && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02final_launches.csv python bench.py --steps 2 --warmup 3 --skip-other > gpurun_out/r02final_ncu_launches.log 2>&1
# full captures: VT f16x3 conv + dense(+head), tiny F=10, tiny F=3
timeout 300 python tools/prof_vt.py f16x3 65536 1 > gpurun_out/r02final_plain_vt.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vt_ -c 2 -f -o gpurun_out/r02_vt_f16x3_v4 python tools/prof_vt.py f16x3 65536 1 > gpurun_out/r02final_ncu_vt.log 2>&1
timeout 300 python tools/prof_small.py tiny10 21 1 > gpurun_out/r02final_plain_tiny.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiny -c 1 -f -o gpurun_out/r02_tiny10_v5 python tools/prof_small.py tiny10 21 1 > gpurun_out/r02final_ncu_tiny.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiny -c 1 -f -o gpurun_out/r02_tiny3_v5 python tools/prof_small.py tiny3 21 1 >> gpurun_out/r02final_ncu_tiny.log 2>&1
tail -n 3 gpurun_out/r02final_pytest_all.log gpurun_out/r02final_smoke.log gpurun_out/r02final_bench.err gpurun_out/r02final_bench_ref.err gpurun_out/r02final_ncu_vt.log gpurun_out/r02final_ncu_tiny.log
