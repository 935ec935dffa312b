#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the Laplace / histogram kernels
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
cat > /tmp/lap.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, datagen
from gpboost_b200 import GPModel
n = int(sys.argv[1])
X, y, _ = datagen.binary_synth(n, 5, False)
gm = GPModel(likelihood="bernoulli_logit", gp_coords=X, gp_approx="vecchia", num_neighbors=30, seed=1)
print(gm.neg_log_likelihood(np.array([1.0, 0.05]), y), gm.laplace_info().tolist())
PY
cat > /tmp/boost.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from gpboost_b200.booster import Booster, Dataset
rng = np.random.default_rng(1); n, F = 1000000, 50
X = rng.random((n, F)); y = 2 * np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.5 * rng.standard_normal(n)
params = dict(objective="regression", num_leaves=31, min_data_in_leaf=20, learning_rate=0.1, max_bin=255, verbose=-1)
b = Booster(params, Dataset(X, y, params=params))
for _ in range(2): b.update()
print("boost done")
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --laplace-n 20000 --boost-n 200000 > gpurun_out/launches_bench.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__cycles_active.avg"
for spec in "trs_fwd_kernel:/tmp/lap.py 100000:3" "trs_bwd_kernel:/tmp/lap.py 100000:3" "v_trs_fwd_kernel:/tmp/lap.py 100000:20"; do
  k="${spec%%:*}"; rest="${spec#*:}"; cmd="${rest%:*}"; skip="${rest##*:}"
  timeout 600 ncu --set full --clock-control none -k "$k" --launch-skip "$skip" -c 1 -f -o gpurun_out/prof_$k python $cmd > gpurun_out/prof_$k.log 2>&1
  echo "== prof_$k.ncu-rep" >> gpurun_out/ncu_raw_summary_r01b.txt
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv --metrics $M 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>=3:
    hdr,units,vals=rows[0],rows[1],rows[-1]
    for h,u,v in zip(hdr,units,vals):
        if h in ('Kernel Name',) or '__' in h: print('   %s = %s %s'%(h,v,u))
" >> gpurun_out/ncu_raw_summary_r01b.txt
done
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/ncu_raw_summary_r01b.txt
wc -l gpurun_out/launches_bench.csv; du -sh gpurun_out
