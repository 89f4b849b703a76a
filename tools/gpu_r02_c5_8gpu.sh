#!/bin/bash
# BASELINE configs[4]: 1e9 frames on 8 GPUs in the <=1e-5 mode (1,907 steps x 65,536 frames x 8 ranks)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 1907 --warmup 3 --skip-other > gpurun_out/r02_bench_c5_8gpu_1e9frames.json 2> gpurun_out/r02_bench_c5_8gpu.err
echo "exit $?" >> gpurun_out/r02_bench_c5_8gpu.err
tail -n 2 gpurun_out/r02_bench_c5_8gpu.err
