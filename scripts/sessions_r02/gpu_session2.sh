#!/bin/bash
# round 2, session 2: full GPU suite with the new defaults + device binning; launch lists of the Laplace evaluation at n=1e6
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s2_pytest.log
cat > /tmp/lap.py <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, datagen
from gpboost_b200 import GPModel
n = int(sys.argv[1])
X, y, _ = datagen.binary_synth(n, 5, False)
gm = GPModel(likelihood="bernoulli_logit", gp_coords=X, gp_approx="vecchia", num_neighbors=30, seed=1)
for rep in range(int(sys.argv[2])):
    t = time.time(); v = gm.neg_log_likelihood(np.array([1.0, 0.05]), y); print(n, time.time() - t, v, gm.laplace_info().tolist(), flush=True)
PY
export GPB200_LAPLACE_TRACE=1
timeout 300 python /tmp/lap.py 1000000 2 2>&1 | tail -4 > gpurun_out/s2_lap_trace.log
unset GPB200_LAPLACE_TRACE
# Newton phase (t = 1 kernels) and SLQ phase (t = 50 kernels): per-launch device times
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 150 --csv --log-file gpurun_out/s2_launches_newton.csv python /tmp/lap.py 1000000 1 > gpurun_out/s2_ncu1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6500 -c 120 --csv --log-file gpurun_out/s2_launches_slq.csv python /tmp/lap.py 1000000 1 > gpurun_out/s2_ncu2.log 2>&1
for f in gpurun_out/s2_pytest.log gpurun_out/s2_lap_trace.log; do echo "== $f"; cat "$f"; done
python - <<'PY'
import csv, collections
for name in ("newton", "slq"):
    try:
        rows = list(csv.reader(l for l in open("gpurun_out/s2_launches_%s.csv" % name) if l.startswith('"')))
    except Exception as e:
        print(name, "no csv", e); continue
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")); u = r[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        k = r[ki].split("(")[0]
        a = agg.setdefault(k, [0, 0.]); a[0] += 1; a[1] += v
    print("==", name)
    for k, (c, s) in agg.items(): print("%-40s n=%4d  avg %9.1f us  total %9.1f us" % (k, c, s / c, s))
PY
