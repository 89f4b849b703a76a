#!/bin/bash
# bf16 fast mode on an 8-GPU box: e2e per input format at N = 8 and N = 1 (the u8 path must scale)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --mode bf16 --skip-other > gpurun_out/r02_bench_bf16_8gpu.json 2> gpurun_out/r02_bench_bf16_8gpu.err
echo "exit $?" >> gpurun_out/r02_bench_bf16_8gpu.err
timeout 600 python bench.py --mode bf16 --skip-other > gpurun_out/r02_bench_bf16_1gpu.json 2> gpurun_out/r02_bench_bf16_1gpu.err
echo "exit $?" >> gpurun_out/r02_bench_bf16_1gpu.err
tail -n 2 gpurun_out/r02_bench_bf16_8gpu.err gpurun_out/r02_bench_bf16_1gpu.err
