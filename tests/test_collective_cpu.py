"""Host-side multi-process logic on CPU (gloo, world_size 2): the row-shard partition and the sum-all-reduce
callback handed to the C++ layer (GPB200_SetCollective)."""
import ctypes
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpboost_b200.parallel import make_allreduce_callback, row_shard
    cb, fn = make_allreduce_callback(dist, device=None)
    buf = (ctypes.c_double * 9)(*[float(rank + 1) * (k + 1) for k in range(9)])
    fn(buf, 9)  # the C++ layer calls it exactly like this
    res = np.array(list(buf))
    rb, re = row_shard(1001, rank, world)
    out.put((rank, res.tolist(), rb, re))
    dist.destroy_process_group()


def test_allreduce_callback_and_row_shards():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    got = sorted([q.get(timeout=120) for _ in range(world)])
    [p.join(60) for p in ps]
    expect = [3. * (k + 1) for k in range(9)]
    for _, res, _, _ in got:
        assert np.allclose(res, expect)
    assert got[0][2] == 0 and got[0][3] == got[1][2] and got[1][3] == 1001
