#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
cat > /tmp/lap.py <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, datagen
from gpboost_b200 import GPModel
n = int(sys.argv[1])
X, y, _ = datagen.binary_synth(n, 5, False)
gm = GPModel(likelihood="bernoulli_logit", gp_coords=X, gp_approx="vecchia", num_neighbors=30, seed=1)
for rep in range(2):
    t = time.time(); v = gm.neg_log_likelihood(np.array([1.0, 0.05]), y); print(n, time.time() - t, v, gm.laplace_info().tolist(), flush=True)
PY
timeout 600 python -m pytest tests/test_laplace_gpu.py -x -q -m gpu 2>&1 | tail -3
export GPB200_LAPLACE_TRACE=1
timeout 300 python /tmp/lap.py 100000 2>&1 | tail -2 | tee gpurun_out/lap_trace.log
timeout 300 python /tmp/lap.py 1000000 2>&1 | tail -2 | tee -a gpurun_out/lap_trace.log
