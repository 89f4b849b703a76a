#!/bin/bash
mkdir -p gpurun_out
python tools/prof_vt.py f16x3 65536 1 > gpurun_out/r02f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vt_ -c 3 -f -o gpurun_out/r02_vt_f16x3_v2 python tools/prof_vt.py f16x3 65536 1 > gpurun_out/r02f_ncu.log 2>&1
tail -5 gpurun_out/r02f_plain.log gpurun_out/r02f_ncu.log
