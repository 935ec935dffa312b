#!/bin/bash
# ncu evidence for the round: launch lists + full captures of the dominant kernels (one GPU, short commands)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --boost-n 0 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vecchia_factor_kernel -s 3 -c 1 -f -o gpurun_out/prof_factor python bench.py --steps 3 --warmup 3 --no-cpu-baseline --boost-n 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:hist_kernel -s 0 -c 1 -f -o gpurun_out/prof_hist python scripts/bench_boost.py 1000000 2000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_tile_kernel -s 2 -c 1 -f -o gpurun_out/prof_syrk python - > /dev/null 2>&1 <<PY
import sys, numpy as np
sys.path.insert(0, "tests"); import datagen
from gpboost_b200 import GPModel
coords, y = datagen.synth(4000, 2, 1)
m = GPModel(gp_coords=coords, cov_function="matern", cov_fct_shape=1.5, gp_approx="none")
print(m.neg_log_likelihood(np.array([0.25, 1.0, 0.1]), y))
PY
ncu --set full --clock-control none -k regex:bt_apply_kernel -c 1 -f -o gpurun_out/prof_btapply python scripts/gpu_first_contact.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
