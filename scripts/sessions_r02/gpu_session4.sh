#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin_reference_package.py -q -m gpu --tb=long 2>&1 | tail -120 | cut -c1-600 > gpurun_out/s4_dropin.log
timeout 900 python -m pytest tests/test_tree_gpu.py tests/test_laplace_gpu.py tests/test_vecchia_gpu.py -q -m gpu --tb=short 2>&1 | tail -60 | cut -c1-400 > gpurun_out/s4_pytest.log
cat gpurun_out/s4_dropin.log gpurun_out/s4_pytest.log
