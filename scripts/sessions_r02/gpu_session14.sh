#!/bin/bash
# round 2, session 14 (1 GPU): (a) tree learner with alternating row-index buffers: parity + timing; (b) likelihood kernel with the
# interleaved covariance rounds / table exp / register-only pivot chain: Vecchia parity tests, headline timing for three group sizes
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecchia_gpu.py tests/test_predict_gpu.py tests/test_tree_gpu.py tests/test_grouped.py tests/test_dropin_reference_package.py -q -m gpu --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/s14_pytest.log
B="--steps 10 --warmup 3 --no-cpu-baseline --boost-n 0 --dense-n 0 --laplace-n 0"
timeout 300 python bench.py $B > gpurun_out/s14_bench_g10.json 2> gpurun_out/s14_bench_g10.err
GPB200_LIB=$PWD/gpboost_b200/lib_gpboost_b200_g6.so timeout 300 python bench.py $B > gpurun_out/s14_bench_g6.json 2> gpurun_out/s14_bench_g6.err
GPB200_LIB=$PWD/gpboost_b200/lib_gpboost_b200_g15.so timeout 300 python bench.py $B > gpurun_out/s14_bench_g15.json 2> gpurun_out/s14_bench_g15.err
timeout 300 python scripts/bench_tree.py 1000000 pingpong: > gpurun_out/s14_tree_bench.log 2>&1
timeout 300 python scripts/mgpu_boost_bench.py 1e6 50 2>&1 | grep "^\[N=" > gpurun_out/s14_boost.log
cat gpurun_out/s14_pytest.log gpurun_out/s14_tree_bench.log gpurun_out/s14_boost.log
python - <<'PY'
import json
for g in ("g10", "g6", "g15"):
    try:
        d = json.loads(open("gpurun_out/s14_bench_%s.json" % g).read().strip().split("\n")[-1])
        print(g, {k: d[k] for k in ("value", "ms_per_step", "negll")}, d["e2e"]["value"], d["roofline_fp64"]["frac"], d["clocks"])
    except Exception as e:
        print(g, "failed", e); print(open("gpurun_out/s14_bench_%s.err" % g).read()[-1500:])
PY
