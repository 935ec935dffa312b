#!/bin/bash
# Laplace-Vecchia bring-up on the GPU box: parity tests + timing at n = 1e5 / 1e6
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_laplace_gpu.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/laplace_pytest.log
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/laplace_timing.log
import sys, time
sys.path.insert(0, 'tests')
import numpy as np, datagen
from gpboost_b200 import GPModel
for n in (100000, 1000000):
    X, y, _ = datagen.binary_synth(n, 5, False)
    t = time.time(); gm = GPModel(likelihood="bernoulli_logit", gp_coords=X, gp_approx="vecchia", num_neighbors=30, seed=1); t1 = time.time()
    v = gm.neg_log_likelihood(np.array([1.0, 0.05]), y); t2 = time.time()
    v2 = gm.neg_log_likelihood(np.array([1.0, 0.05]), y); t3 = time.time()
    print(n, 'create %.2fs first eval %.3fs second eval %.3fs' % (t1 - t, t2 - t1, t3 - t2), repr(v), gm.laplace_info().tolist(), flush=True)
PY
