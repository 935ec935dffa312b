#!/bin/bash
# round 2, session 15 (1 GPU): ncu --set full of the reworked likelihood kernel (source page for the stall distribution)
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vecchia_nll2_kernel -s 3 -c 1 -f -o gpurun_out/s15_prof_nll2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --boost-n 0 --laplace-n 0 --dense-n 0 > gpurun_out/s15_ncu_nll2.log 2>&1
tail -3 gpurun_out/s15_ncu_nll2.log; ls -la gpurun_out/s15_prof_nll2.ncu-rep
