#!/bin/bash
# round 2, session 10 (1 GPU): the round's bench command (both arms) + ncu evidence for profiles/
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
# launch list of the same command (headline part: first 400 launches after the model creation)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s10_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --laplace-n 0 --dense-n 0 > gpurun_out/s10_ncu_launches.log 2>&1
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vecchia_nll2_kernel -s 3 -c 1 -f -o gpurun_out/s10_prof_nll2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --boost-n 0 --laplace-n 0 --dense-n 0 > gpurun_out/s10_ncu_nll2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:hist3_kernel -s 0 -c 1 -f -o gpurun_out/s10_prof_hist3 python scripts/bench_tree.py 1000000 default: > gpurun_out/s10_ncu_hist3.log 2>&1
for r in nll2 hist3; do
  ncu -i gpurun_out/s10_prof_$r.ncu-rep --page raw --csv 2>/dev/null | python - "$r" <<'PY' >> gpurun_out/s10_ncu_summary.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem")
    print("== prof_%s" % sys.argv[1])
    for h, u, v in zip(hdr, units, vals):
        if h in keep: print("   %s = %s %s" % (h, v, u))
PY
done
ls -la gpurun_out/s10_prof_*.ncu-rep 2>/dev/null
cat gpurun_out/s10_ncu_summary.txt
tail -3 gpurun_out/s10_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/s10_bench.json").read().strip().split("\n")[-1])
    def short(o, depth=0):
        if isinstance(o, dict): return {k: short(v, depth + 1) for k, v in o.items() if k not in ("note", "sample", "call", "peak_source", "traffic_unit")}
        return o
    print(json.dumps(short(d), indent=1)[:6000])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/s10_bench.err").read()[-2000:])
PY
# reference arm (CPU on the box's host cores; the driver runs it at round end)
( time timeout 1200 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/s10_bench_reference.json 2> gpurun_out/s10_bench_reference.err
tail -c 2500 gpurun_out/s10_bench_reference.json; tail -4 gpurun_out/s10_bench_reference.err
