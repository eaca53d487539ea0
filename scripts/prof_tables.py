#!/usr/bin/env python
"""Wall time of one table build (reo_tables) against the number of reference columns, on a staged 30k x 20k matrix: how
long is one round of tile pairs, what does a launch cost on top.  python scripts/prof_tables.py [ncols ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
r, n = 30000, 10000
t, group, _ = pkg.synth.scrna_torch(r, n, n, device="cuda:0")
_, gid = pkg.api.group_levels(group)
h = pkg.Reo(0, seed=7)
if os.environ.get("REO_FAKE_WORLD"):
    h.set_collective(0, int(os.environ["REO_FAKE_WORLD"]), lambda ptr, nbytes: None)
dm = pkg.DeviceMatrix(t.data_ptr(), pkg._lib.REO_I64, r, 2 * n, r, keepalive=t)
h.stage(dm, gid, 2)
sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256, 448, 1024, 3408]
rng = np.random.default_rng(0)
for nc in sizes:
    mask = np.zeros(r, bool)
    mask[rng.choice(r, nc, replace=False)] = True
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.tables(0, mask)
        best = min(best, time.perf_counter() - t0)
    tp = 469 * ((nc + 63) // 64 + 1) // 2
    print(f"ncols {nc:6d}: {best * 1e3:8.3f} ms   (~{tp} tile pairs of 64 x 128 x 20000, {tp / 296:.2f} per CTA)")
