#!/bin/bash
# round 2, session 25 (1 GPU): the gradient pass also writes the factor, a store request at that state is answered without a pass:
# parity (Vecchia, trees / GPBoost goldens, drop-in package), GPBoost iteration time with and without the shortcut
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecchia_gpu.py tests/test_predict_gpu.py tests/test_tree_gpu.py tests/test_dropin_reference_package.py -q -m gpu --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/s25_pytest.log
timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s25_boost.log
GPB200_GRAD_STORES=0 timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" | sed 's/^/[GRAD_STORES=0] /' >> gpurun_out/s25_boost.log
timeout 200 python scripts/time_vecchia_modes.py > gpurun_out/s25_modes.log 2>&1
cat gpurun_out/s25_pytest.log gpurun_out/s25_boost.log gpurun_out/s25_modes.log
