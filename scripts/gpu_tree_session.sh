#!/bin/bash
# GPU session for the tree learner: parity (default, then every opt-in tree kernel), timings, ncu of hist2_kernel / split scan, launch list.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_default.log
GPB200_HIST_KERNEL=2 GPB200_PARTITION=2 timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_hist2.log
GPB200_HIST_KERNEL=2 GPB200_PARTITION=2 GPB200_FUSED_SCAN=1 timeout 300 python -m pytest tests/test_tree_gpu.py tests/test_grouped.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_fused.log
timeout 300 python scripts/bench_tree.py 1000000 hist1:GPB200_HIST_KERNEL=1 hist2:GPB200_HIST_KERNEL=2 hist2_part2:GPB200_HIST_KERNEL=2,GPB200_PARTITION=2 \
  hist2_part2_fused:GPB200_HIST_KERNEL=2,GPB200_PARTITION=2,GPB200_FUSED_SCAN=1 devloop:GPB200_HIST_KERNEL=2,GPB200_PARTITION=2,GPB200_FUSED_SCAN=1,GPB200_TREE_LOOP=device graph:GPB200_HIST_KERNEL=2,GPB200_PARTITION=2,GPB200_FUSED_SCAN=1,GPB200_TREE_LOOP=graph > gpurun_out/bench_tree.log 2>&1
GPB200_HIST_KERNEL=2 timeout 240 ncu --set full --clock-control none --import-source on -k regex:hist2_kernel -c 2 -o gpurun_out/hist2 -f \
  python scripts/bench_tree.py 1000000 hist2:GPB200_HIST_KERNEL=2 > gpurun_out/ncu_hist2.log 2>&1
GPB200_HIST_KERNEL=2 timeout 240 ncu --set full --clock-control none --import-source on -k regex:split_scan_kernel -s 20 -c 1 -o gpurun_out/scan -f \
  python scripts/bench_tree.py 200000 hist2:GPB200_HIST_KERNEL=2 > gpurun_out/ncu_scan.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tree_hist2.csv \
  python scripts/bench_tree.py 1000000 hist2_part2_fused:GPB200_HIST_KERNEL=2,GPB200_PARTITION=2,GPB200_FUSED_SCAN=1 > gpurun_out/ncu_tree2.log 2>&1
GPB200_TREE_LOOP=device timeout 300 python -m pytest tests/test_tree_gpu.py tests/test_grouped.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_devloop.log
GPB200_TREE_LOOP=graph timeout 300 python -m pytest tests/test_tree_gpu.py tests/test_grouped.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_graph.log
cat gpurun_out/pytest_gpu_graph.log gpurun_out/pytest_gpu_devloop.log gpurun_out/pytest_gpu_default.log gpurun_out/pytest_gpu_hist2.log gpurun_out/pytest_gpu_fused.log gpurun_out/bench_tree.log
