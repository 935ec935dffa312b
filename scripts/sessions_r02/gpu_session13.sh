#!/bin/bash
# round 2, session 13 (1 GPU): arg-max + selector + next planner fused into one launch: tree parity, timing against the three-launch form
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tree_gpu.py tests/test_grouped.py tests/test_dropin_reference_package.py -q -m gpu --tb=short 2>&1 | tail -20 | cut -c1-300 > gpurun_out/s13_pytest.log
timeout 300 python scripts/bench_tree.py 1000000 fused_advance: three_launches:GPB200_FUSED_ADVANCE=0 > gpurun_out/s13_tree_bench.log 2>&1
cat gpurun_out/s13_pytest.log gpurun_out/s13_tree_bench.log
