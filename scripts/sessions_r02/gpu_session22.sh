#!/bin/bash
# round 2, session 22 (1 GPU): warm-cache per-kernel times of the tree learner (ncu --cache-control none: L2 keeps the chunk partials
# between the histogram kernel and the scan, as in the graph replay)
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
GPB200_TREE_LOOP=device timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 700 --csv --log-file gpurun_out/s22_launches_tree_warm.csv python scripts/bench_tree.py 1000000 default: > gpurun_out/s22_ncu.log 2>&1
tail -3 gpurun_out/s22_ncu.log; wc -l gpurun_out/s22_launches_tree_warm.csv
