#!/usr/bin/env python
"""One identify_degs job on a device-resident synthetic matrix, for ncu / timing.
python scripts/prof_job.py <workload> [reps] [n_iter]      (workload names: bench.py WORKLOADS)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

pkg = ge.load_package()
wl = sys.argv[1] if len(sys.argv) > 1 else "c5_allref_30kx20k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 128
kind, r, n1, n2, n_ref, _ = bench.WORKLOADS[wl]
if kind == "scrna":
    t, group, is_de = pkg.synth.scrna_torch(r, n1, n2, device="cuda:0")
else:
    data, group, is_de = pkg.synth.bulk(r, n1, n2)
    t = torch.from_numpy(np.ascontiguousarray(data.T)).to("cuda:0")
ref = pkg.synth.reference_mask(is_de, n_ref) if n_ref > 0 else np.ones(r, dtype=bool)
_, gid = pkg.api.group_levels(group)
h = pkg.Reo(0, seed=pkg.synth.TIE_SEED)
if os.environ.get("REO_FAKE_WORLD"):
    # timing study on ONE GPU of what rank 0 of a `world`-rank job does: the pair-tile space is partitioned for `world`
    # ranks and the (dummy) collective gathers nothing, so the tables are partial and the RESULTS ARE WRONG -- only the
    # launch times mean anything
    h.set_collective(0, int(os.environ["REO_FAKE_WORLD"]), lambda ptr, nbytes: None)
dm = pkg.DeviceMatrix(t.data_ptr(), pkg._lib.REO_I64, r, n1 + n2, r, keepalive=t)
for i in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = h.identify_degs(dm, gid, 2, ref, 0.01, 1.0, 0.05, n_iter, 5)
    dt = time.perf_counter() - t0
    s = out.stats
    print(f"{wl} rep {i}: wall {dt * 1e3:.2f} ms  stage {s['ms_stage']:.3f} pairs {s['ms_pairs']:.3f} stats {s['ms_stats']:.3f} "
          f"total {s['ms_total']:.3f}  evals {s['iters_done']} n_deg {s['n_deg']} B {s['rank_bits']} W {s['sample_words']} "
          f"cmp {s['compares']:.4e} launches {s['pair_launches']}/{s['kernel_launches']}")
