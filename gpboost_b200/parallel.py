"""Multi-GPU plumbing: one process per GPU (torchrun), observations row-sharded, NCCL only on shard boundaries.

The C++ layer never talks to NCCL itself: like the reference's LGBM_NetworkInitWithFunctions
(include/LightGBM/c_api.h:1306-1317) it accepts the collective as a function pointer. Here that function is a
ctypes callback around `torch.distributed.all_reduce` (backend nccl over NVLink on the GPU box, gloo in CPU tests).
What crosses ranks per likelihood evaluation is 9 fp64 sums."""
import ctypes

import numpy as np

ALLREDUCE_CFUNC = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.c_int)


def row_shard(n, rank, world):
    """[begin, end) of ordered observations owned by `rank` — must equal the partition in REModel's constructor
    (gpboost_b200/csrc/host/re_model.cpp)."""
    chunk = (n + world - 1) // world
    b = min(n, chunk * rank)
    return b, min(n, b + chunk)


def make_allreduce_callback(dist, device=None):
    """Returns (ctypes callback object to keep alive, python callable). `device`: torch device for NCCL, None for gloo."""
    import torch

    def _allreduce(buf, count):
        arr = np.ctypeslib.as_array(buf, shape=(count,))
        t = torch.from_numpy(arr.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        arr[:] = t.cpu().numpy()

    return ALLREDUCE_CFUNC(_allreduce), _allreduce


def init_collective(lib, dist, device=None):
    """Register this process' rank / world size and the all-reduce with the C++ layer. Keep the returned object alive."""
    cb, _ = make_allreduce_callback(dist, device)
    rc = lib.GPB200_SetCollective(dist.get_rank(), dist.get_world_size(), ctypes.cast(cb, ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(lib.LGBM_GetLastError().decode())
    return cb


def init_nccl(lib, dist, device_index):
    """Native collective: the C++ runtime opens its own NCCL communicator (csrc/host/collective.cpp) and all-reduces device
    buffers on the engines' streams. torch.distributed is only the launcher-side channel that broadcasts the 128-byte id."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    if lib.GPB200_SetDevice(int(device_index)) != 0:
        raise RuntimeError(lib.LGBM_GetLastError().decode())
    buf = ctypes.create_string_buffer(128)
    if rank == 0 and lib.GPB200_NcclGetUniqueId(buf) != 0:
        raise RuntimeError(lib.LGBM_GetLastError().decode())
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.to(torch.device("cuda", int(device_index)))
    dist.broadcast(t, src=0)
    ident = bytes(t.cpu().numpy().tobytes())
    if lib.GPB200_NcclInit(rank, world, ctypes.c_char_p(ident)) != 0:
        raise RuntimeError(lib.LGBM_GetLastError().decode())
