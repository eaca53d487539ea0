"""One process per GPU.  The pair kernel's tile space -- symmetric sweep or rows x columns -- is cut into work
items that are dealt round-robin to the ranks (every world-th item, scrambled inside each supertile: csrc/reo_pairs2.cu);
every rank therefore holds partial sums for ALL genes, the per-gene Int32 tables are summed over the ranks (NCCL
all-reduce inside the library, or any all-gather handed in through reo_set_collective followed by a local sum) and K3-K6
run replicated on every rank (deterministic FP64 -> identical masks, one collective per evaluation, no mask broadcast).
K1 is sharded by sample words when the matrix is large.  Nothing here re-implements the partition: `pair_plan` asks the
library itself (reo_debug_pair_plan, host code, no GPU needed) which tile pairs a rank evaluates."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

TILE = 64


def pair_plan(r: int, ncols: int, sample_words: int, planes: int, rank: int, world: int, one_sided: bool = False):
    """The library's own partition for a table build over `ncols` of `r` genes: (triples, nsym_tiles, ntr, ntc) with
    triples[:, 0] = row tile, [:, 1] = column tile (panel coordinates), [:, 2] = 1 (pairs update their row genes) |
    2 (... their column genes).  nsym_tiles > 0: the first nsym_tiles row tiles are the column panel itself."""
    lib = L.load()
    ns, ntr, ntc = C.c_int32(), C.c_int32(), C.c_int32()
    args = (int(r), int(ncols), int(sample_words), int(planes), int(rank), int(world), 1 if one_sided else 0)
    n = lib.reo_debug_pair_plan(*args, None, 0, C.byref(ns), C.byref(ntr), C.byref(ntc))
    if n < 0:
        raise ValueError("reo_debug_pair_plan: bad argument")
    out = np.zeros((max(int(n), 1), 3), dtype=np.int32)
    lib.reo_debug_pair_plan(*args, out.ctypes.data, int(n), C.byref(ns), C.byref(ntr), C.byref(ntc))
    return out[:int(n)], ns.value, ntr.value, ntc.value


class DevBuf:
    """A raw device pointer exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def make_torch_allgather(rank: int, world: int, group=None):
    """In-place all-gather callback for Reo.set_collective, over torch.distributed (NCCL on GPUs): the library gathers
    the ranks' partial tables with it and sums them itself."""
    import torch
    import torch.distributed as dist

    def allgather(dev_ptr: int, bytes_per_rank: int):
        full = torch.as_tensor(DevBuf(dev_ptr, bytes_per_rank * world), device="cuda")
        mine = full[rank * bytes_per_rank:(rank + 1) * bytes_per_rank].clone()
        dist.all_gather_into_tensor(full, mine, group=group)
        torch.cuda.current_stream().synchronize()

    return allgather


def init_nccl_in_library(handle, rank: int, world: int, group=None):
    """One process per GPU: create the library's own NCCL communicator (tables are then summed on the library's stream
    with no host round trip).  The 128-byte unique id travels over torch.distributed."""
    import torch.distributed as dist
    from . import api
    box = [api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    handle.comm_init(rank, world, box[0])
