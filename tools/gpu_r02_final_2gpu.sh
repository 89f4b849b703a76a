#!/bin/bash
# 2-GPU box: the full GPU suite (incl. the two-device test), the async soak, and the 2-GPU bench
mkdir -p gpurun_out
rm -f gpurun_out/r02q_*
timeout 2400 python -m pytest tests -x -q -m gpu -rs > gpurun_out/r02q_pytest_all.log 2>&1
echo "exit $?" >> gpurun_out/r02q_pytest_all.log
timeout 600 python tools/stress_async.py > gpurun_out/r02q_stress_async.log 2>&1
echo "exit $?" >> gpurun_out/r02q_stress_async.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/r02q_bench_2gpu.json 2> gpurun_out/r02q_bench_2gpu.err
echo "exit $?" >> gpurun_out/r02q_bench_2gpu.err
tail -n 6 gpurun_out/r02q_pytest_all.log gpurun_out/r02q_stress_async.log gpurun_out/r02q_bench_2gpu.err
