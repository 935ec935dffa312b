#!/bin/bash
# round 2, session 12 (1 GPU): full suite on the final code (STORE pass in the two-observation layout, sliced y upload), smoke, bench
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -25 | cut -c1-300 > gpurun_out/s12_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s12_smoke.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err
timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s12_boost.log
cat gpurun_out/s12_pytest.log; tail -3 gpurun_out/s12_smoke.log; cat gpurun_out/s12_boost.log; tail -4 gpurun_out/s12_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s12_bench.json").read().strip().split("\n")[-1])
print({k: d[k] for k in ("value", "ms_per_step", "negll", "gpu_launches")}, d["e2e"]["value"], d["e2e"].get("pageable_y"), d["roofline_fp64"]["frac"], d["clocks"])
for k in ("gpboost", "gpboost_grouped", "laplace", "dense"):
    v = d.get(k, {})
    print(k, {a: b for a, b in v.items() if a in ("iters_per_sec", "ms_per_iter", "evals_per_sec", "sec_per_eval", "trees_only_ms_per_iter")})
PY
