"""CPU suite: the N>1 path (row-tile sharding + table all-gather) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp
from conftest import ROOT, small_case


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import __graft_entry__ as ge
    from importlib import import_module
    pkg = ge.load_package()
    oracle, co = ge.load_oracle()
    d = import_module(pkg.__name__ + ".dist")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data, group = small_case(3, 300, 9, 11)
    levels, gid = oracle.group_levels(group)
    thr = co.thresholds_for(gid, 2, 0.01)
    mask = np.arange(300) % 4 != 0
    nt = -(-300 // 64)
    tpr, ranges = d.shard_plan(nt, world)
    t0, t1 = ranges[rank]
    i0, i1 = min(300, t0 * 64), min(300, t1 * 64)
    local, _ = co.block_tables(data, gid, 2, thr, np.nonzero(mask)[0], seed=7, i0=i0, i1=i1) if i1 > i0 else (np.zeros((0, 9)), 0)
    full = d.allgather_rows_cpu(local, rank, world, nt)[:300]
    want, _ = co.block_tables(data, gid, 2, thr, np.nonzero(mask)[0], seed=7)
    ok = bool(np.array_equal(full, want)) and d.table_slice_bytes(nt, world) == tpr * 64 * 36
    ret[rank] = ok
    dist.destroy_process_group()


def test_row_tile_sharding_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
