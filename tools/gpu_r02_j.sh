#!/bin/bash
# 8-GPU box: scaling of value and of the e2e variants
mkdir -p gpurun_out
for N in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02j_bench_${N}gpu.json 2> gpurun_out/r02j_bench_${N}gpu.err
  echo "exit $?" >> gpurun_out/r02j_bench_${N}gpu.err
done
nvidia-smi topo -m > gpurun_out/r02j_topo.txt 2>&1
lscpu | head -25 > gpurun_out/r02j_lscpu.txt 2>&1
tail -n 3 gpurun_out/r02j_bench_8gpu.err gpurun_out/r02j_bench_2gpu.err
