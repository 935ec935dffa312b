#!/bin/bash
# round 2, session 27 (1 GPU): the round's bench line on the final code
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/s27_bench.json 2> gpurun_out/s27_bench.err
tail -4 gpurun_out/s27_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s27_bench.json").read().strip().split("\n")[-1])
print({k: d[k] for k in ("value", "ms_per_step", "negll", "gpu_launches")}, d["e2e"]["value"], d["roofline_fp64"]["frac"], d["clocks"])
for k in ("gpboost", "gpboost_grouped", "laplace", "dense", "cpu_baseline"):
    v = d.get(k, {})
    print(k, {a: b for a, b in v.items() if a in ("iters_per_sec", "ms_per_iter", "evals_per_sec", "sec_per_eval", "trees_only_ms_per_iter", "value", "cores")})
PY
