#!/bin/bash
# round 2, session 23 (1 GPU): final code — full GPU suite, smoke, the round's bench line, per-mode pass times, launch list and
# ncu --set full captures of the likelihood and gradient passes (profiles/r02_*)
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -25 | cut -c1-300 > gpurun_out/s23_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s23_smoke.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/s23_bench.json 2> gpurun_out/s23_bench.err
timeout 200 python scripts/time_vecchia_modes.py > gpurun_out/s23_modes.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s23_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --laplace-n 0 --dense-n 0 > gpurun_out/s23_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vecchia_nll2_kernel -s 3 -c 1 -f -o gpurun_out/s23_prof_nll2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --boost-n 0 --laplace-n 0 --dense-n 0 > gpurun_out/s23_ncu_nll2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:nll2_kernel<.*2>' -s 2 -c 1 -f -o gpurun_out/s23_prof_grad python scripts/time_vecchia_modes.py > gpurun_out/s23_ncu_grad.log 2>&1
: > gpurun_out/s23_ncu_summary.txt
for r in nll2 grad; do
  ncu -i gpurun_out/s23_prof_$r.ncu-rep --page raw --csv 2>/dev/null | python - "$r" <<'PY' >> gpurun_out/s23_ncu_summary.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio")
    print("== prof_%s" % sys.argv[1])
    for h, u, v in zip(hdr, units, vals):
        if h in keep: print("   %s = %s %s" % (h, v, u))
PY
done
cat gpurun_out/s23_pytest.log; tail -3 gpurun_out/s23_smoke.log; cat gpurun_out/s23_modes.log; cat gpurun_out/s23_ncu_summary.txt; tail -4 gpurun_out/s23_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s23_bench.json").read().strip().split("\n")[-1])
print({k: d[k] for k in ("value", "ms_per_step", "negll", "gpu_launches")}, d["e2e"]["value"], d["e2e"].get("pageable_y", {}).get("value"), d["roofline_fp64"]["frac"], d["clocks"])
for k in ("gpboost", "gpboost_grouped", "laplace", "dense", "cpu_baseline"):
    v = d.get(k, {})
    print(k, {a: b for a, b in v.items() if a in ("iters_per_sec", "ms_per_iter", "evals_per_sec", "sec_per_eval", "trees_only_ms_per_iter", "value", "cores")})
PY
