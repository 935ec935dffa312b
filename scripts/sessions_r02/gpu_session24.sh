#!/bin/bash
# round 2, session 24 (1 GPU): the evidence files of session 23 again (its gpurun_out exceeded the 64 MiB return limit): bench line,
# launch list, ncu --set full of the likelihood and the gradient pass (only the likelihood report travels back)
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/s24_bench.json 2> gpurun_out/s24_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s24_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --laplace-n 0 --dense-n 0 > gpurun_out/s24_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vecchia_nll2_kernel -s 3 -c 1 -f -o gpurun_out/s24_prof_nll2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --boost-n 0 --laplace-n 0 --dense-n 0 > gpurun_out/s24_ncu_nll2.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:nll2_kernel<.*2>' -s 2 -c 1 -f -o gpurun_out/s24_prof_grad python scripts/time_vecchia_modes.py > gpurun_out/s24_ncu_grad.log 2>&1
: > gpurun_out/s24_ncu_summary.txt
for r in nll2 grad; do
  ncu -i gpurun_out/s24_prof_$r.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_raw_summary.py $r >> gpurun_out/s24_ncu_summary.txt
done
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/s24_prof_grad.ncu-rep
cat gpurun_out/s24_ncu_summary.txt; tail -3 gpurun_out/s24_ncu_grad.log; du -sh gpurun_out
