#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 450 --csv --log-file gpurun_out/launches_tree.csv python scripts/bench_boost.py 1000000 2000 > gpurun_out/ncu_tree.log 2>&1
tail -3 gpurun_out/ncu_tree.log
