#!/bin/bash
# round 2, session 7 (1 GPU): tiled operator kernels (bulk-copy staging): parity + timings; head-size sweep of the t=1 solves
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_laplace_gpu.py -q -m gpu --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/s7_pytest.log
cat > /tmp/lap.py <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, datagen
from gpboost_b200 import GPModel
n = int(sys.argv[1])
X, y, _ = datagen.binary_synth(n, 5, False)
t0 = time.time()
gm = GPModel(likelihood="bernoulli_logit", gp_coords=X, gp_approx="vecchia", num_neighbors=30, seed=1)
print("create", time.time() - t0, flush=True)
for rep in range(int(sys.argv[2])):
    t = time.time(); v = gm.neg_log_likelihood(np.array([1.0, 0.05]), y); print(n, time.time() - t, v, gm.laplace_info().tolist(), flush=True)
PY
export GPB200_LAPLACE_TRACE=1
for v in "default:" "untiled:GPB200_LAPLACE_TILED=0" "head64:GPB200_TRS_VARIANT=1,GPB200_TRS_HEAD_ROWS=64" "head1024:GPB200_TRS_VARIANT=1,GPB200_TRS_HEAD_ROWS=1024" "head4096:GPB200_TRS_VARIANT=1,GPB200_TRS_HEAD_ROWS=4096"; do
  name=${v%%:*}; envs=${v#*:}; envs=${envs//,/ }
  echo "== $name" >> gpurun_out/s7_lap_trace.log
  env $envs timeout 200 python /tmp/lap.py 1000000 2 2>&1 | tail -4 >> gpurun_out/s7_lap_trace.log
done
unset GPB200_LAPLACE_TRACE
# launch list of the SLQ phase with the tiled kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6500 -c 120 --csv --log-file gpurun_out/s7_launches_slq.csv python /tmp/lap.py 1000000 1 > gpurun_out/s7_ncu.log 2>&1
python - <<'PY'
import csv, collections
try:
    rows = list(csv.reader(l for l in open("gpurun_out/s7_launches_slq.csv") if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")); u = r[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        k = r[ki].split("(")[0]; a = agg.setdefault(k, [0, 0.]); a[0] += 1; a[1] += v
    for k, (c, s) in agg.items(): print("%-40s n=%4d  avg %9.1f us" % (k, c, s / c))
except Exception as e:
    print("no csv", e)
PY
for f in gpurun_out/s7_pytest.log gpurun_out/s7_lap_trace.log; do echo "== $f"; cat $f; done
