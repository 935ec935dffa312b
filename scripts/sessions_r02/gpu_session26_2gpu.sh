#!/bin/bash
# round 2, session 26 (2 GPUs): multi-GPU parity and the boosting iteration at N=2 on the final code (gradient pass stores the factor,
# response not re-installed between the optimiser and the gradient call, line search)
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -25 | cut -c1-600 > gpurun_out/s26_mgpu.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s26_boost_n2.log
cat gpurun_out/s26_mgpu.log gpurun_out/s26_boost_n2.log
