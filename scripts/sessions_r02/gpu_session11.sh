#!/bin/bash
# round 2, session 11 (1 GPU): gradient pass with two observations per warp: parity (GP fits, GPBoost goldens), timing of the boosting iteration
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecchia_gpu.py tests/test_tree_gpu.py tests/test_dropin_reference_package.py tests/test_predict_gpu.py -q -m gpu --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/s11_pytest.log
timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s11_boost_grad2.log
GPB200_NLL_KERNEL=1 timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s11_boost_grad1.log
timeout 300 python - > gpurun_out/s11_fit.log 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from gpboost_b200 import GPModel
rng = np.random.default_rng(1)
n = 1000000
coords = rng.random((n, 2)); y = np.sin(5 * coords[:, 0]) * np.cos(4 * coords[:, 1]) + 0.5 * rng.standard_normal(n)
m = GPModel(gp_coords=coords, cov_function="matern", cov_fct_shape=1.5, gp_approx="vecchia", num_neighbors=30, vecchia_ordering="random", seed=1)
t = time.time(); m.fit(y); print("fit n=1e6: %.3f s, %d iterations, cov pars %s, negll %.6f" % (time.time() - t, m._get_num_optim_iter(), m.get_cov_pars(), m.get_current_neg_log_likelihood()))
PY
for f in gpurun_out/s11_*.log; do echo "== $f"; cat $f; done
