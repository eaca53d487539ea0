#!/usr/bin/env python
"""K1 only (staging of a device-resident single-cell matrix), for ncu / timing.  python scripts/prof_stage.py [r] [cells_per_group] [reps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
r = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t, group, _ = pkg.synth.scrna_torch(r, n, n, device="cuda:0")
_, gid = pkg.api.group_levels(group)
h = pkg.Reo(0, seed=7)
dm = pkg.DeviceMatrix(t.data_ptr(), pkg._lib.REO_I64, r, 2 * n, r, keepalive=t)
for i in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = h.stage(dm, gid, 2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gb = r * 2 * n * 8 / 1e9
    print(f"stage {i}: {dt * 1e3:.3f} ms  ({gb / dt:.0f} GB/s of Int64 input)  {info}")
