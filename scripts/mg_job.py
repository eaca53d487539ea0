#!/usr/bin/env python
"""torchrun helper: one identify_degs job per rank set (NCCL inside the library), REO_TIMING breakdown on stderr.
torchrun --nproc-per-node N scripts/mg_job.py <workload> [reps]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

pkg = ge.load_package()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = sys.argv[1] if len(sys.argv) > 1 else "c5_allref_30kx20k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
data, dev, gid, ref = bench.make_workload(pkg, wl, device=f"cuda:{local}")
kind, r, n1, n2, n_ref, _ = bench.WORKLOADS[wl]
if dev is None:
    dev = torch.from_numpy(np.ascontiguousarray(data.T)).to(f"cuda:{local}")
h = pkg.Reo(local, seed=pkg.synth.TIE_SEED)
if world > 1:
    from importlib import import_module
    import_module(pkg.__name__ + ".dist").init_nccl_in_library(h, rank, world)
dm = pkg.DeviceMatrix(dev.data_ptr(), pkg._lib.REO_I64, r, n1 + n2, r, keepalive=dev)
for i in range(reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if i == reps - 1:
        os.environ["REO_TIMING"] = "1"
    out = h.identify_degs(dm, gid, 2, ref, 0.01, 1.0, 0.05, 128, 5)
    s = out.stats
    if rank == 0:
        print(f"rep {i}: total {s['ms_total']:.3f} stage {s['ms_stage']:.3f} pairs {s['ms_pairs']:.3f} other {s['ms_stats']:.3f}", flush=True)
h.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
