#!/bin/bash
# round 2, sessions 19-20 (1 GPU): likelihood / store pass with the next pair's covariance rounds inside the factorisation (vecchia_nll3):
# parity tests, headline timing against the plain two-observation kernel (GPB200_NLL_KERNEL=2), GPBoost iteration
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecchia_gpu.py tests/test_predict_gpu.py -q -m gpu --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/s19_pytest.log
B="--steps 10 --warmup 3 --no-cpu-baseline --boost-n 0 --dense-n 0 --laplace-n 0"
timeout 300 python bench.py $B > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err
GPB200_NLL_KERNEL=2 timeout 300 python bench.py $B > gpurun_out/s19_bench_nll2.json 2> gpurun_out/s19_bench_nll2.err
timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s19_boost.log
cat gpurun_out/s19_pytest.log gpurun_out/s19_boost.log
python - <<'PY'
import json
for g in ("", "_nll2"):
    try:
        d = json.loads(open("gpurun_out/s19_bench%s.json" % g).read().strip().split("\n")[-1])
        print(g, {k: d[k] for k in ("value", "ms_per_step", "negll")}, d["e2e"]["value"], d["roofline_fp64"]["frac"], d["clocks"])
    except Exception as e:
        print(g, "failed", e); print(open("gpurun_out/s19_bench%s.err" % g).read()[-1500:])
PY
