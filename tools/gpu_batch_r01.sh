set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest6.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke6.log 2>&1
python bench.py > gpurun_out/bench_bf16_v10.json 2> gpurun_out/bench_bf16_v10.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-other > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vt_ -c 2 -f -o gpurun_out/r01_vt_bf16_v10 python tools/prof_vt.py bf16 65536 1 > gpurun_out/ncu_vt_v10.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vt_ -c 3 -f -o gpurun_out/r01_vt_tf32x3 python tools/prof_vt.py tf32x3 18944 1 > gpurun_out/ncu_vt_tf32.log 2>&1
cat gpurun_out/pytest6.log gpurun_out/smoke6.log
