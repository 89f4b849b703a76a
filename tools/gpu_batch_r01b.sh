set -x
ncu --set full --clock-control none --import-source on -k regex:vt_ -c 2 -f -o gpurun_out/r01_vt_bf16_final python tools/prof_vt.py bf16 65536 1 > gpurun_out/ncu_vt_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vt_ -c 3 -f -o gpurun_out/r01_vt_tf32x3_final python tools/prof_vt.py tf32x3 18944 1 > gpurun_out/ncu_vt_tf32_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"q612|tiny|fwht" -c 4 -f -o gpurun_out/r01_small_final python tools/prof_small.py all 21 1 > gpurun_out/ncu_small_final.log 2>&1
ncu --set full --clock-control none -k regex:sdr -c 1 -f -o gpurun_out/r01_sdr_final python tools/prof_sdr.py 26 > gpurun_out/ncu_sdr_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --skip-other > gpurun_out/ncu_bench_final.log 2>&1
ls -la gpurun_out/*final*
