"""Row-sharded dense QP over the GPUs of one box (include/qpalm_b200.h part 2b, csrc/shard.cu).

One process per GPU.  ``init(rank, world)`` creates the library's NCCL communicator: rank 0 makes the id, the ranks
exchange it through ``torch.distributed`` (plumbing; any broadcast works), then every rank calls the ordinary
``Qpalm("b200")`` API with the same full problem and gets the same full solution back."""
from __future__ import annotations

import ctypes as C

from .interface import load_library


def init(rank: int, world: int, device=None) -> None:
    import torch
    import torch.distributed as dist
    lib = load_library("b200")
    lib.qpalm_b200_shard_unique_id.argtypes = [C.c_char_p]
    lib.qpalm_b200_shard_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
    if world <= 1:
        return
    buf = C.create_string_buffer(128)
    if rank == 0 and lib.qpalm_b200_shard_unique_id(buf):
        raise RuntimeError("ncclGetUniqueId failed")
    t = torch.tensor(list(buf.raw), dtype=torch.uint8)
    if device is not None and dist.get_backend() == "nccl":
        t = t.to(device)
    dist.broadcast(t, src=0)
    ident = bytes(t.cpu().tolist())
    if lib.qpalm_b200_shard_init(rank, world, ident):
        raise RuntimeError("qpalm_b200_shard_init failed")


def finalize() -> None:
    load_library("b200").qpalm_b200_shard_finalize()


def row_block(m: int, rank: int, world: int):
    """Constraint rows [lo, lo + cnt) kept by `rank`, and the padded block size `cap` of the in-place allgather --
    the partition engine_create (csrc/kernels.cu) applies when qpalm_b200_shard_init was called: equal blocks of
    cap = ceil(m / world) rows, the last ranks possibly short or empty."""
    cap = -(-m // world)
    lo = min(rank * cap, m)
    cnt = max(0, min(cap, m - rank * cap))
    return lo, cnt, cap
