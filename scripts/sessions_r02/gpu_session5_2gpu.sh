#!/bin/bash
# two GPUs: multi-GPU parity with the device-resident data-parallel leaf loop (default) and with the host loop; boosting timings
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -25 | cut -c1-600 > gpurun_out/s5_mgpu_device.log
GPB200_SHARDED_LOOP=host timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-400 > gpurun_out/s5_mgpu_host.log
timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 > gpurun_out/s5_boost_n1.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/mgpu_boost_bench.py 1e6 50 > gpurun_out/s5_boost_n2_device.log 2>&1
GPB200_SHARDED_LOOP=host timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/mgpu_boost_bench.py 1e6 50 > gpurun_out/s5_boost_n2_host.log 2>&1
timeout 600 python -m pytest tests/test_dropin_reference_package.py tests/test_vecchia_gpu.py -q -m gpu --tb=short 2>&1 | tail -15 | cut -c1-400 > gpurun_out/s5_dropin.log
for f in gpurun_out/s5_*.log; do echo "== $f"; grep -v "^$" "$f" | tail -25; done
