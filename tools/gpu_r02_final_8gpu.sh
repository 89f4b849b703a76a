#!/bin/bash
# 8-GPU box: scaling of value and of the e2e variants at N = 8, 4 and 2
mkdir -p gpurun_out
rm -f gpurun_out/r02x_*
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N > gpurun_out/r02x_bench_${N}gpu.json 2> gpurun_out/r02x_bench_${N}gpu.err
  echo "exit $?" >> gpurun_out/r02x_bench_${N}gpu.err
done
tail -n 2 gpurun_out/r02x_bench_8gpu.err gpurun_out/r02x_bench_4gpu.err gpurun_out/r02x_bench_2gpu.err
