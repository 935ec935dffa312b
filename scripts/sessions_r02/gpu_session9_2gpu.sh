#!/bin/bash
# two GPUs, tight timeouts: data-parallel leaf loop enqueued eagerly on the device (default) — parity, then timings
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -25 | cut -c1-600 > gpurun_out/s9_mgpu_device.log
if grep -q "1 passed" gpurun_out/s9_mgpu_device.log; then
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s9_boost_n2_device.log
  GPB200_SHARDED_LOOP=host timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s9_boost_n2_host.log
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s9_bench_2gpu.json 2> gpurun_out/s9_bench_2gpu.err
fi
for f in gpurun_out/s9_*.log; do echo "== $f"; cat $f; done
tail -c 3000 gpurun_out/s9_bench_2gpu.json 2>/dev/null
tail -5 gpurun_out/s9_bench_2gpu.err 2>/dev/null
