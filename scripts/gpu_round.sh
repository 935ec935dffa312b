#!/bin/bash
# Round check on the GPU box: GPU parity tests, smoke, both bench arms. Logs -> gpurun_out/
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 --laplace-ref-n 100000 2>gpurun_out/bench_ref.err | tail -1 | tee gpurun_out/bench_reference.json
timeout 900 python bench.py 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
