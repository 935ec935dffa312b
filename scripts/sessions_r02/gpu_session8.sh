#!/bin/bash
# round 2, session 8 (1 GPU): two-observations-per-warp likelihood kernel: parity + timing against the one-observation kernel; full suite
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecchia_gpu.py tests/test_laplace_gpu.py -q -m gpu --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/s8_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --boost-n 0 --laplace-n 0 --dense-n 0 > gpurun_out/s8_bench_nll2.json 2> gpurun_out/s8_bench_nll2.err
GPB200_NLL_KERNEL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --boost-n 0 --laplace-n 0 --dense-n 0 > gpurun_out/s8_bench_nll1.json 2> gpurun_out/s8_bench_nll1.err
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -15 | cut -c1-300 > gpurun_out/s8_pytest_all.log
for f in gpurun_out/s8_pytest.log gpurun_out/s8_pytest_all.log; do echo "== $f"; cat $f; done
for f in gpurun_out/s8_bench_nll2 gpurun_out/s8_bench_nll1; do echo "== $f"; python -c "
import json,sys
try:
    d=json.loads(open('$f.json').read().strip().split('\n')[-1]); print({k:d[k] for k in ('value','ms_per_step','negll','gpu_launches')}, d['e2e']['value'], d['roofline_fp64']['frac'], d['clocks'])
except Exception as e:
    print('fail', e); print(open('$f.err').read()[-1500:])
"; done
