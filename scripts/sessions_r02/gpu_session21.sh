#!/bin/bash
# round 2, session 21 (1 GPU): gradient pass with the derivative pair values parked in the upper triangle (three CTAs per SM):
# parity tests, per-mode pass times for the three builds, GPBoost iteration
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecchia_gpu.py tests/test_predict_gpu.py -q -m gpu --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/s21_pytest.log
timeout 200 python scripts/time_vecchia_modes.py > gpurun_out/s21_modes.log 2>&1
GPB200_LIB=$PWD/gpboost_b200/lib_var_grad2.so timeout 200 python scripts/time_vecchia_modes.py >> gpurun_out/s21_modes.log 2>&1
GPB200_LIB=$PWD/gpboost_b200/lib_var_gradreg.so timeout 200 python scripts/time_vecchia_modes.py >> gpurun_out/s21_modes.log 2>&1
timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s21_boost.log
cat gpurun_out/s21_pytest.log gpurun_out/s21_modes.log gpurun_out/s21_boost.log
