"""Row-tile sharding across one process per GPU.  The pair kernel's work is independent per
gene row, so gene-row tiles are split into `world` contiguous, equal-sized slices (the last ranks may own
fewer or no real tiles) and the per-gene Int32 tables are all-gathered; K3-K6 then run replicated on every
rank (deterministic FP64 -> identical masks, one collective per evaluation, no mask broadcast)."""
from __future__ import annotations

TILE = 64


def shard_plan(n_tiles: int, world: int):
    """(tiles_per_rank, [(t0, t1) per rank]) -- mirrors launch_tables() in csrc/reo_api.cu."""
    tpr = -(-n_tiles // world)
    return tpr, [(min(n_tiles, q * tpr), min(n_tiles, (q + 1) * tpr)) for q in range(world)]


def word_shard_plan(n_words: int, world: int):
    """K1 sharding: (words_per_rank, [(w_lo, w_hi) per rank]) over the staged sample words -- mirrors do_stage() in
    csrc/reo_api.cu.  Ranks past the last word stage nothing; the all-gathered blocks are padded to words_per_rank."""
    wq = -(-n_words // world)
    return wq, [(min(n_words, q * wq), min(n_words, (q + 1) * wq)) for q in range(world)]


def k1_is_sharded(r: int, c: int, world: int, data_on_device: bool) -> bool:
    """Staging is sharded from 2^24 values resident in HBM, from 2^20 values for host input (mirrors do_stage())."""
    return world > 1 and r * c >= (1 << 24 if data_on_device else 1 << 20)


def table_slice_bytes(n_tiles: int, world: int) -> int:
    tpr, _ = shard_plan(n_tiles, world)
    return tpr * TILE * 9 * 4


class DevBuf:
    """A raw device pointer exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def make_torch_allgather(rank: int, world: int, group=None):
    """In-place all-gather callback for Reo.set_collective, over torch.distributed (NCCL on GPUs)."""
    import torch
    import torch.distributed as dist

    def allgather(dev_ptr: int, bytes_per_rank: int):
        full = torch.as_tensor(DevBuf(dev_ptr, bytes_per_rank * world), device="cuda")
        mine = full[rank * bytes_per_rank:(rank + 1) * bytes_per_rank].clone()
        dist.all_gather_into_tensor(full, mine, group=group)
        torch.cuda.current_stream().synchronize()

    return allgather


def allgather_rows_cpu(local_rows, rank: int, world: int, n_tiles: int, group=None):
    """Host-side twin of the table exchange (gloo): every rank contributes the [tpr*64, 9] Int32 slice of
    the tiles it owns and receives the full table.  Used by the CPU tests of the N>1 path."""
    import torch
    import torch.distributed as dist
    tpr, _ = shard_plan(n_tiles, world)
    mine = torch.zeros((tpr * TILE, 9), dtype=torch.int32)
    mine[: local_rows.shape[0]] = torch.as_tensor(local_rows, dtype=torch.int32)
    full = torch.zeros((world * tpr * TILE, 9), dtype=torch.int32)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full.numpy()


def init_nccl_in_library(handle, rank: int, world: int, group=None):
    """One process per GPU: create the library's own NCCL communicator (tables are then all-gathered on the
    library's stream with no host round trip).  The 128-byte unique id travels over torch.distributed."""
    import torch.distributed as dist
    from . import api
    box = [api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    handle.comm_init(rank, world, box[0])
