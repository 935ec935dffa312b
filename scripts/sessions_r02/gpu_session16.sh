#!/bin/bash
# round 2, session 16 (1 GPU): likelihood kernel with the two-deep gather pipeline, the column store after the rank-1 updates and the
# pivot taken from per-lane diagonal registers: Vecchia parity tests, headline timing, GPBoost iteration
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecchia_gpu.py tests/test_predict_gpu.py -q -m gpu --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/s16_pytest.log
B="--steps 10 --warmup 3 --no-cpu-baseline --boost-n 0 --dense-n 0 --laplace-n 0"
timeout 300 python bench.py $B > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err
timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s16_boost.log
cat gpurun_out/s16_pytest.log gpurun_out/s16_boost.log
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/s16_bench.json").read().strip().split("\n")[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "negll")}, d["e2e"]["value"], d["roofline_fp64"]["frac"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s16_bench.err").read()[-1500:])
PY
