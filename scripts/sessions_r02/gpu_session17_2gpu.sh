#!/bin/bash
# round 2, session 17 (2 GPUs): final code — multi-GPU parity (sharded Vecchia rows with the sliced y upload, eager device-resident
# data-parallel tree loop with alternating index buffers, Laplace probes), then the bench line at N=2 including the configs[3] shape
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -25 | cut -c1-600 > gpurun_out/s17_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --config3 on > gpurun_out/s17_bench_2gpu.json 2> gpurun_out/s17_bench_2gpu.err
cat gpurun_out/s17_mgpu.log
tail -c 6000 gpurun_out/s17_bench_2gpu.json 2>/dev/null
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/s17_bench_2gpu.err | tail -8
