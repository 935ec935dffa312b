"""Timing of data-parallel boosting under torchrun (one process per GPU): plain L2 boosting and GPBoost-Vecchia iterations at n x F.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/mgpu_boost_bench.py [n] [F]
Prints ms/iter on rank 0 and a hash of the model text (identical trees over N and leaf loops show as identical hashes)."""
import hashlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from gpboost_b200 import GPModel, load_lib
from gpboost_b200.booster import Booster, Dataset
from gpboost_b200.parallel import init_nccl

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
lib = load_lib()
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world)
    init_nccl(lib, dist, local)
assert lib.GPB200_SetDevice(local) == 0
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = np.random.default_rng(1)
X = rng.random((n, F)).astype(np.float32 if n * F > 2e8 else np.float64)
coords = rng.random((n, 2))
y = 2 * np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + np.sin(5 * coords[:, 0]) * np.cos(4 * coords[:, 1]) + 0.5 * rng.standard_normal(n)
params = dict(objective="regression", num_leaves=31, min_data_in_leaf=20, learning_rate=0.1, max_bin=255, verbose=-1)
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
t0 = time.perf_counter(); ds = Dataset(X, y, params=params); t_ds = time.perf_counter() - t0
b = Booster(params, ds)
b.update(); sync()
t0 = time.perf_counter()
for _ in range(10): b.update()
sync(); dt = (time.perf_counter() - t0) / 10
h = hashlib.sha1(b.model_to_string().encode()).hexdigest()[:12]
if rank == 0: print("[N=%d loop=%s] plain boosting n=%d F=%d: dataset %.2fs, %.3f ms/iter (%.1f iters/s) model %s" % (world, os.environ.get("GPB200_SHARDED_LOOP", "device"), n, F, t_ds, dt * 1e3, 1 / dt, h), flush=True)
del b
gp = GPModel(gp_coords=coords, cov_function="matern", cov_fct_shape=1.5, gp_approx="vecchia", num_neighbors=30, vecchia_ordering="random", seed=1)
b = Booster(params, ds, gp_model=gp)
t0 = time.perf_counter(); b.update(); sync(); first = time.perf_counter() - t0
t0 = time.perf_counter()
for _ in range(5): b.update()
sync(); dt = (time.perf_counter() - t0) / 5
if rank == 0: print("[N=%d loop=%s] GPBoost Vecchia m=30 n=%d F=%d: first %.2fs, %.3f ms/iter (%.1f iters/s) cov_pars %s" % (world, os.environ.get("GPB200_SHARDED_LOOP", "device"), n, F, first, dt * 1e3, 1 / dt, gp.get_cov_pars()), flush=True)
if world > 1:
    dist.barrier(); lib.GPB200_NcclFinalize(); dist.destroy_process_group()
