#!/bin/bash
# round 2, session 3: full GPU suite (binning, prediction, m <= 60, drop-in with the unmodified reference package)
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/s3_pytest.log
cat gpurun_out/s3_pytest.log
