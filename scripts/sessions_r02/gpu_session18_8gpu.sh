#!/bin/bash
# round 2, session 18 (8 GPUs): the bench line at N=8 (configs[3] shape included: n=5e6 x 100 features, rows sharded over 8 ranks)
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s18_bench_8gpu.json 2> gpurun_out/s18_bench_8gpu.err
tail -c 7000 gpurun_out/s18_bench_8gpu.json 2>/dev/null
grep -v "^\*\|OMP_NUM\|^$\|UserWarning\|return func" gpurun_out/s18_bench_8gpu.err | tail -8
