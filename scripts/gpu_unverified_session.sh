#!/bin/bash
# First GPU session after round 1: everything that was written after the round's GPU budget was spent, in order of dependency.
# Each step is a separate process (a CUDA fault in one must not hide the others) with its own timeout; logs -> gpurun_out/.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
# 0. the defaults (must stay green)
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/u0_default.log
# 1. factor kernel MODE_STORE_GRAD against the oracle (A, D^-1, dA, dD of the latent factor)
GPB200_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_laplace_gpu.py -q -m gpu -k latent_factor_range_derivative 2>&1 | tail -15 > gpurun_out/u1_store_grad.log
# 2. Laplace gradient (iterative branch) against the reference's gradients
GPB200_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests/test_laplace_gpu.py -q -m gpu -k laplace_gradient_matches 2>&1 | tail -25 > gpurun_out/u2_laplace_grad.log
# 3. the fit built on it
GPB200_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_laplace_gpu.py -q -m gpu -k laplace_fit_matches 2>&1 | tail -25 > gpurun_out/u3_laplace_fit.log
# 4. hist3_kernel (padded tile, on-demand peer gradients): parity through the variant test's cases, then timing against hist2
GPB200_HIST_KERNEL=3 timeout 300 python -m pytest tests/test_tree_gpu.py tests/test_grouped.py -x -q -m gpu -k "not variants" 2>&1 | tail -6 > gpurun_out/u4_hist3_parity.log
timeout 300 python scripts/bench_tree.py 1000000 hist2:GPB200_HIST_KERNEL=2 hist3:GPB200_HIST_KERNEL=3 > gpurun_out/u4_hist3_bench.log 2>&1
# 5. reduce_scan2_kernel (one threshold per thread)
GPB200_FUSED_SCAN=2 timeout 300 python -m pytest tests/test_tree_gpu.py tests/test_grouped.py -x -q -m gpu -k "not variants" 2>&1 | tail -6 > gpurun_out/u5_scan2_parity.log
timeout 300 python scripts/bench_tree.py 1000000 scan1:GPB200_FUSED_SCAN=1 scan2:GPB200_FUSED_SCAN=2 scan2_hist3:GPB200_FUSED_SCAN=2,GPB200_HIST_KERNEL=3 > gpurun_out/u5_scan2_bench.log 2>&1
for f in gpurun_out/u*.log; do echo "== $f"; cat "$f"; done
