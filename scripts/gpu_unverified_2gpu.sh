#!/bin/bash
# Two-GPU session (gpurun --gpus 2): the multi-GPU parity check with the current defaults, then with the two-kernel partition
# enabled for row-sharded learners. Logs -> gpurun_out/.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/m0_default.log
GPB200_PARTITION_SHARDED=2 timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/m1_partition2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
  > gpurun_out/m2_bench_2gpu.json 2> gpurun_out/m2_bench_2gpu.err
for f in gpurun_out/m*.log gpurun_out/m2_bench_2gpu.json; do echo "== $f"; tail -5 "$f"; done
