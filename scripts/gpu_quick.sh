#!/bin/bash
# quick GPU check: parity tests + bench (no ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python scripts/gpu_first_contact.py > gpurun_out/first_contact.log 2>&1
