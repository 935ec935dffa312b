#!/bin/bash
# round 2, session 6 (1 GPU): full suite with the new defaults; Laplace n=1e6 timings per solve / order variant; histogram variants
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40 | cut -c1-300 > gpurun_out/s6_pytest.log
cat > /tmp/lap.py <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, datagen
from gpboost_b200 import GPModel
n = int(sys.argv[1])
X, y, _ = datagen.binary_synth(n, 5, False)
gm = GPModel(likelihood="bernoulli_logit", gp_coords=X, gp_approx="vecchia", num_neighbors=30, seed=1)
for rep in range(int(sys.argv[2])):
    t = time.time(); v = gm.neg_log_likelihood(np.array([1.0, 0.05]), y); print(n, time.time() - t, v, gm.laplace_info().tolist(), flush=True)
PY
export GPB200_LAPLACE_TRACE=1
for v in "default:" "trs0:GPB200_TRS_VARIANT=0" "index_order:GPB200_LAPLACE_ORDER=index" "sleep0:GPB200_TRS_SLEEP_NS=0" "sleep20:GPB200_TRS_SLEEP_NS=20" "ctas8:GPB200_TRS_CTAS_PER_SM=8" "ctas3:GPB200_TRS_CTAS_PER_SM=3"; do
  name=${v%%:*}; envs=${v#*:}
  echo "== $name" >> gpurun_out/s6_lap_trace.log
  env $envs timeout 200 python /tmp/lap.py 1000000 2 2>&1 | tail -3 >> gpurun_out/s6_lap_trace.log
done
unset GPB200_LAPLACE_TRACE
timeout 300 python scripts/bench_tree.py 1000000 hist3_red:GPB200_HIST_KERNEL=3 hist3_plain:GPB200_HIST_KERNEL=4 hist2:GPB200_HIST_KERNEL=2 > gpurun_out/s6_hist_bench.log 2>&1
for f in gpurun_out/s6_pytest.log gpurun_out/s6_lap_trace.log gpurun_out/s6_hist_bench.log; do echo "== $f"; cat $f; done
