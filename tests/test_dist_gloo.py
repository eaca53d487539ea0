"""CPU suite: the N>1 path with world_size 2 over gloo.  Every rank asks the LIBRARY which tile pairs it evaluates
(reo_debug_pair_plan: the partition code of csrc/reo_pairs2.cu replayed on the host), computes exactly those with the
CPU oracle, the partial tables are summed over gloo (the all-reduce the GPU path does over NCCL) and compared with the
oracle's tables -- for the symmetric all-genes sweep, the permuted [C ; N] panel and the one-sided shape."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp
from conftest import ROOT, small_case

TILE = 64


def _partial_tables(d, co, data, gid, thr, cols, rank, world, r, one_sided):
    """Partial r x 9 tables of `rank`: the oracle's categories, accumulated only over the rank's tile pairs."""
    W, NP = 1, 9                                   # 9 + 11 samples share one word; 8 rank bits + coin plane
    trip, nsym, ntr, ntc = d.pair_plan(r, len(cols), W, NP, rank, world, one_sided=one_sided)
    in_c = np.zeros(r, bool)
    in_c[cols] = True
    if nsym > 0 and len(cols) < r:                 # permuted panel [C ascending, pad | N ascending]
        n_tiles_n = -(-int((~in_c).sum()) // TILE)
        nsymp = ntr - n_tiles_n
        row_gene = -np.ones(ntr * TILE, dtype=np.int64)
        row_gene[:len(cols)] = cols
        rest = np.nonzero(~in_c)[0]
        row_gene[nsymp * TILE:nsymp * TILE + len(rest)] = rest
        col_gene = row_gene
    elif nsym > 0:                                 # every gene is a column: the staged planes themselves
        row_gene = -np.ones((ntr + 1) * TILE, dtype=np.int64)
        row_gene[:r] = np.arange(r)
        col_gene = row_gene
    else:                                          # one-sided: all genes x gathered columns
        row_gene = -np.ones(ntr * TILE, dtype=np.int64)
        row_gene[:r] = np.arange(r)
        col_gene = -np.ones((ntc + 1) * TILE, dtype=np.int64)
        col_gene[:len(cols)] = cols
    cat = co.categories(data, gid, 2, thr, seed=7)          # [r, r] categories 1..9, 0 on the diagonal
    tab = np.zeros((r, 9), dtype=np.int64)
    seen = set()
    for I, J, fl in trip:
        assert (I, J) not in seen
        seen.add((int(I), int(J)))
        gi = row_gene[I * TILE:(I + 1) * TILE]
        gj = col_gene[J * TILE:(J + 1) * TILE]
        gi, gj = gi[gi >= 0], gj[gj >= 0]
        for a in gi:
            for b in gj:
                if a == b:
                    continue
                q = int(cat[a, b])
                if fl & 1:
                    tab[a, q - 1] += 1
                if fl & 2:
                    tab[b, 10 - q - 1] += 1          # the mirrored category for the column gene (src:385-386)
    return tab, len(trip)


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from importlib import import_module
    pkg = ge.load_package()
    oracle, co = ge.load_oracle()
    d = import_module(pkg.__name__ + ".dist")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok, sizes = True, []
    r = 1100
    data, group = small_case(3, r, 9, 11)
    levels, gid = oracle.group_levels(group)
    thr = co.thresholds_for(gid, 2, 0.01)
    for cols, one_sided in ((np.arange(r), False),                          # symmetric sweep over all genes
                            (np.nonzero(np.arange(r) % 13 != 0)[0], False),  # >= 1024 columns: permuted [C ; N] panel
                            (np.nonzero(np.arange(r) % 4 == 0)[0], False),   # few columns: one-sided
                            (np.arange(r), True)):                           # all columns, forced one-sided
        tab, n = _partial_tables(d, co, data, gid, thr, cols, rank, world, r, one_sided)
        t = torch.from_numpy(tab)
        dist.all_reduce(t)                                                    # what ncclAllReduce does on the GPUs
        want, _ = co.block_tables(data, gid, 2, thr, cols, seed=7)
        ok = ok and bool(np.array_equal(t.numpy(), want))
        sizes.append(n)
    ret[rank] = (ok, sizes)
    dist.destroy_process_group()


def test_pair_partition_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0][0] and ret[1][0]
    assert all(a > 0 and b > 0 for a, b in zip(ret[0][1], ret[1][1]))     # both ranks worked in every shape
