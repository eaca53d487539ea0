#!/usr/bin/env python
"""
bench.py -- the REO hot path (identify_degs, src/RankCompV3.jl:339-438) on BASELINE.json's metric:
gene-pair x sample comparisons per second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libreo_cuda.so)
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference (oracle/)

A "step" is one complete identify_degs job over one synthetic expression matrix.  Default workload = BASELINE.json
configs[4], the shape north_star names: 30 000 genes x 20 000 cells (10k per group), every gene a reference gene,
n_iter = 128 -- 30k x 30k gene pairs x 20k cells per table build.
  value : executed comparisons / t with the matrix already resident in HBM (raw Int64, Julia layout) -- staging,
          every pair-kernel launch, every statistics kernel and the result read-back are inside the timed region.
  e2e   : the same job through the public call with HOST buffers (H2D of the 4.8 GB matrix inside the timed
          region): `pinned` input (the headline e2e) and `pageable` input (what julia/reo_ccall.jl hands over).
Comparisons are counted as EXECUTED (gene pair, sample) evaluations: the pair kernel uses the reference's mirror
property (src:385-386), so an all-genes build evaluates r(r-1)/2 pairs, exactly the reference's own is_greater call
count (src:366-372), not r^2.
N > 1 (torchrun, one process per GPU): supertiles of the pair-tile space are dealt round-robin to the ranks, per-gene
tables are summed over NCCL once per evaluation -- the total work is fixed, i.e. strong scaling.
Every run checks itself against the CPU oracle (oracle/, test infrastructure) AFTER the timed region: a 24-row block
of tables bit-exact, the job's own tables against the row-sum invariant, and at N > 1 the whole result against a
single-rank run.  A mismatch exits non-zero.
The Julia reference cannot run in this image (no julia binary), so --impl reference and cpu_baseline time the plain-C
restatement oracle/reo_oracle.c ("port") on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

METRIC = "gene_pair_sample_comparisons_per_sec"
UNIT = "comparisons/s"

WORKLOADS = {
    # name: (kind, genes, n1, n2, n_ref (0 = all genes), description)
    "c5_allref_30kx20k": ("scrna", 30000, 10000, 10000, 0, "BASELINE configs[4]: 30k x 30k gene pairs x 20k cells, all genes as references"),
    "c4_scrna_30kx20k": ("scrna", 30000, 10000, 10000, 3000, "BASELINE configs[3]: 30k genes x 10k vs 10k cells, 3000 initial references"),
    "c2_bulk_20kx200": ("bulk", 20000, 100, 100, 3000, "BASELINE configs[1]: bulk 20k genes x 100 vs 100, 3000 HK refs"),
    "c3_pseudobulk_25kx100": ("bulk", 25000, 50, 50, 3000, "BASELINE configs[2] core shape: 25k genes x 50 vs 50"),
    "mid_scrna_8kx4k": ("scrna", 8000, 2000, 2000, 1500, "mid-size single-cell: 8k genes x 2k vs 2k cells"),
    "tiny": ("bulk", 2000, 20, 20, 300, "smoke-sized"),
}
JOB = dict(pval_reo=0.01, pval_deg=1.0, padj_deg=0.05, n_iter=128, n_conv=5)


def make_workload(pkg, name, device=None, host=True):
    """Returns (column-major r x c Int64 numpy matrix or None, device tensor [c, r] or None, group ids, reference mask).
    Single-cell shapes are generated on the GPU when one is given (same model, seconds instead of minutes)."""
    kind, r, n1, n2, n_ref, _ = WORKLOADS[name]
    dev = None
    if kind == "bulk":
        data, group, is_de = pkg.synth.bulk(r, n1, n2)
        data = np.asfortranarray(data)
    elif device is not None:
        dev, group, is_de = pkg.synth.scrna_torch(r, n1, n2, device=device)
        data = None
    else:
        data, group, is_de = pkg.synth.scrna(r, n1, n2)
        data = np.asfortranarray(data)
    ref = pkg.synth.reference_mask(is_de, n_ref) if n_ref > 0 else np.ones(r, dtype=bool)
    _, gid = pkg.api.group_levels(group)
    return data, dev, gid, ref


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev = dev
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # samples under load only: idle samples sit at the base clock
        hi = [x for x, p in zip(sm, pw) if p >= 0.5 * max(pw)] if sm else []
        return {"sm_mhz": statistics.median(hi) if hi else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm),
                "samples_under_load": len(hi)}


def numpy_scrna_sub(pkg, genes, n1, n2, seed=1234):
    """A `genes` x (n1 + n2) block of the single-cell model (same generator as synth.scrna, fewer genes): the CPU arm
    times the oracle on it -- a bounded sample of the workload."""
    data, group, _ = pkg.synth.scrna(genes, n1, n2, seed=seed)
    _, gid = pkg.api.group_levels(group)
    return data, gid


def cpu_block(co, pkg, data, gid, rows, cols):
    """One timed call of the plain-C restatement: tables of `rows` gene rows against `cols` reference genes."""
    thr = co.thresholds_for(gid, 2, JOB["pval_reo"])
    t0 = time.perf_counter()
    _, n = co.block_tables(data, gid, 2, thr, cols, seed=pkg.synth.TIE_SEED, i0=0, i1=rows)
    return n, time.perf_counter() - t0


def cpu_baseline(pkg, co, name, target_s=15.0):
    """Bounded sample of the workload on the host cores: a 256-row block of gene rows against the reference columns
    of a sub-matrix with the workload's samples, scaled to ~target_s of CPU work."""
    kind, r, n1, n2, n_ref, _ = WORKLOADS[name]
    if kind == "scrna":
        gsub = 2048
        data, gid = numpy_scrna_sub(pkg, gsub, n1, n2)
        cols = np.arange(gsub, dtype=np.int32)
    else:
        data, _, gid, ref = make_workload(pkg, name)
        cols = np.nonzero(ref)[0].astype(np.int32)
        gsub = data.shape[0]
    d = np.asfortranarray(data.astype(np.float64))
    rows = 32
    n, dt = cpu_block(co, pkg, d, gid, rows, cols)       # also pays the oracle's transposition once
    rows = int(max(32, min(gsub, 256, rows * target_s / max(dt, 1e-3))))
    n, dt = cpu_block(co, pkg, d, gid, rows, cols)
    c = n1 + n2
    return dict(value=n / dt, unit=UNIT, cores=co.num_threads(), kind="port",
                sample=f"{rows} gene rows x {len(cols)} reference genes x {c} samples = {n:.3e} comparisons in {dt:.2f} s "
                       f"on a {gsub}-gene block of the workload (oracle/reo_oracle.c, OpenMP; the Julia reference cannot "
                       f"run here); full job = extrapolation in rows"), (d, gid, rows, cols)


def run_reference(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, co = ge.load_oracle()
    co.use_all_cores()                      # torchrun exports OMP_NUM_THREADS=1
    base, (d, gid, rows, cols) = cpu_baseline(pkg, co, args.workload, target_s=max(1.0, 45.0 / max(args.steps + args.warmup, 1)))
    times, n = [], 0
    for s in range(args.warmup + args.steps):
        n, dt = cpu_block(co, pkg, d, gid, rows, cols)
        if s >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    val = n / t
    base["value"] = val
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "note": WORKLOADS[args.workload][5],
                   "sample": f"each step = {rows} gene rows x {len(cols)} refs x {d.shape[1]} samples (bounded sample of the workload)"},
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def result_digest(out):
    """Hash of everything a caller receives that must not depend on the number of GPUs."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(out.result).tobytes())
    h.update(np.ascontiguousarray(out.updown).tobytes())
    h.update(np.ascontiguousarray(out.final_ref).tobytes())
    h.update(np.asarray(out.iters, dtype=np.int64).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5_allref_30kx20k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary measurements (small bulk job, pseudo-bulk from cells, bundled data, float path)")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-input e2e measurement")
    ap.add_argument("--collective", default="nccl", choices=["nccl", "torch"],
                    help="N>1: table reduction by NCCL inside the library (default) or through torch.distributed")
    args = ap.parse_args()
    pkg = ge.load_package()
    if args.impl == "reference":
        return run_reference(args, pkg)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the REO path has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)

    kind, r, n1, n2, n_ref, note = WORKLOADS[args.workload]
    c = n1 + n2
    data, dev, gid, ref = make_workload(pkg, args.workload, device=device)
    h = pkg.Reo(local, seed=pkg.synth.TIE_SEED)
    if world > 1:
        from importlib import import_module
        dmod = import_module(pkg.__name__ + ".dist")
        if args.collective == "nccl":
            dmod.init_nccl_in_library(h, rank, world)
        else:
            h.set_collective(rank, world, dmod.make_torch_allgather(rank, world))

    # inputs: an HBM-resident copy (value), a pinned host copy (e2e) and a plain pageable host copy (e2e.pageable);
    # all column-major r x c Int64 = Julia's Matrix{Int64}
    if dev is None:
        dev = torch.from_numpy(np.ascontiguousarray(data.T)).to(device)      # [c, r] row-major == r x c column-major
    host = torch.empty((c, r), dtype=torch.int64, pin_memory=True)
    host.copy_(dev)
    dmat = pkg.DeviceMatrix(dev.data_ptr(), pkg._lib.REO_I64, r, c, r, keepalive=dev)
    hmat = host.numpy().T                                                    # F-ordered view of the pinned buffer
    nbytes = r * c * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)         # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def one(mat):
        return h.identify_degs(mat, gid, 2, ref, JOB["pval_reo"], JOB["pval_deg"], JOB["padj_deg"], JOB["n_iter"], JOB["n_conv"])

    def timed(mat, steps, warmup):
        """K steps, each bracketed by barrier + synchronize on both sides, device time by CUDA events, MAX over ranks."""
        out = None
        for _ in range(warmup):
            out = None
            out = one(mat)
        total_ms, stats = 0.0, []
        by_rank = [0.0] * world
        for _ in range(steps):
            flush.fill_(1)                      # L2 flush between timed iterations
            out = None                          # release the previous result buffers before the next call
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = one(mat)                      # blocking: returns after the result read-back
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=device)
            if dist is not None:
                every = [torch.zeros_like(ms) for _ in range(world)]
                dist.all_gather(every, ms)
                for i, t in enumerate(every):
                    by_rank[i] += float(t.item()) / steps
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            else:
                by_rank[0] += float(ms.item()) / steps
            total_ms += float(ms.item())
            stats.append(out.stats)
        return total_ms, stats, out, [round(x, 4) for x in by_rank]

    def whole_job(stats, key):
        t = torch.tensor([float(sum(s[key] for s in stats))], device=device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t)
        return t.item()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, st_dev, out_dev, ms_by_rank = timed(dmat, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, st_e2e, out_e2e, _ = timed(hmat, args.steps, 1)
    e2e_pageable = None
    if not args.no_pageable:
        pg = np.empty((c, r), dtype=np.int64)               # plain malloc'ed memory, as a Julia Matrix would be
        np.copyto(pg, host.numpy())
        steps_pg = max(2, min(args.steps, 5))
        ms_pg, st_pg, out_pg, _ = timed(pg.T, steps_pg, 1)
        cmp_pg = whole_job(st_pg, "compares")
        e2e_pageable = {"value": cmp_pg / (ms_pg * 1e-3), "unit": UNIT, "ms_per_step": ms_pg / steps_pg, "steps": steps_pg,
                        "note": "input in pageable host memory (what julia/reo_ccall.jl passes): the library stages it through "
                                "pinned bounce buffers"}
        assert result_digest(out_pg) == result_digest(out_dev), "pageable-input result differs from the device-resident one"
        del pg, out_pg
    # every rank's own view of its last timed step (pair kernels, staging, everything else incl. waiting in collectives)
    mine = torch.tensor([st_dev[-1]["ms_pairs"], st_dev[-1]["ms_stage"], st_dev[-1]["ms_stats"], st_dev[-1]["ms_total"]],
                        device=device, dtype=torch.float64)
    per_rank = [mine.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(per_rank, mine)
    per_rank = [[round(float(v), 3) for v in t.tolist()] for t in per_rank]
    cmp_dev = whole_job(st_dev, "compares")
    cmp_e2e = whole_job(st_e2e, "compares")
    launches = whole_job(st_dev, "kernel_launches")
    assert result_digest(out_e2e) == result_digest(out_dev), "host-input result differs from the device-resident one"

    # ---- parity, after the timed region (oracle = test infrastructure; never inside a timed call) -------------------
    _, co = ge.load_oracle()
    co.use_all_cores()
    parity = {"rows": 24, "cols": 4096}
    res = out_dev.result[0]
    tab = res[:, 2:11].astype(np.int64)
    fr = out_dev.final_ref[0].astype(bool)
    parity["row_sum_invariant"] = bool(np.array_equal(tab.sum(axis=1), fr.sum() - fr.astype(np.int64)))
    rng = np.random.default_rng(0)
    rows_g = np.sort(rng.choice(r, 24, replace=False))
    cols_g = np.sort(rng.choice(np.nonzero(fr)[0], min(4096, int(fr.sum())), replace=False))
    genes = np.union1d(rows_g, cols_g)
    sub = dev[:, torch.as_tensor(genes, device=device)].cpu().numpy().T.astype(np.int64)      # [genes, c]
    thr = co.thresholds_for(gid, 2, JOB["pval_reo"])
    mask = np.zeros(r, bool)
    mask[cols_g] = True
    got = h.tables(0, mask, thresholds=thr)          # collective at N > 1; the staged matrix is still resident
    if rank == 0:
        want = co.block_tables_idx(sub, genes, gid, 2, thr, np.searchsorted(genes, rows_g), np.searchsorted(genes, cols_g),
                                   seed=pkg.synth.TIE_SEED)
        parity["tables_bit_exact"] = bool(np.array_equal(got[rows_g], want))
        se_w, p_w = co.empirical_null(res[:, 11])
        rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))  # noqa: E731
        parity["pval_max_rel_err"] = rel(res[:, 0], p_w)
        parity["padj_max_rel_err"] = rel(res[:, 1], co.bh(p_w))
        parity["pvals_within_1e-12"] = bool(parity["pval_max_rel_err"] <= 1e-12 and parity["padj_max_rel_err"] <= 1e-12)
    parity["vs_single_rank"] = None
    if world > 1:
        # the whole result must not depend on the number of ranks: rank 0 repeats the job alone on its GPU
        ok = torch.ones(1, device=device)
        if rank == 0:
            with pkg.Reo(local, seed=pkg.synth.TIE_SEED) as h1:
                o1 = h1.identify_degs(dmat, gid, 2, ref, JOB["pval_reo"], JOB["pval_deg"], JOB["padj_deg"], JOB["n_iter"], JOB["n_conv"])
            parity["vs_single_rank"] = {
                "tables": bool(np.array_equal(o1.result[:, :, 2:11], out_dev.result[:, :, 2:11])),
                "updown": bool(np.array_equal(o1.updown, out_dev.updown)),
                "final_ref": bool(np.array_equal(o1.final_ref, out_dev.final_ref)),
                "iters": o1.iters == out_dev.iters,
                "everything_bitwise": result_digest(o1) == result_digest(out_dev)}
            ok[0] = float(all(parity["vs_single_rank"].values()))
        digests = [None] * world
        dist.all_gather_object(digests, result_digest(out_dev))
        parity["all_ranks_same_result"] = len(set(digests)) == 1
        dist.broadcast(ok, src=0)
    parity_ok = True
    if rank == 0:
        parity_ok = bool(parity["row_sum_invariant"] and parity["tables_bit_exact"] and parity["pvals_within_1e-12"]
                         and (world == 1 or (all(parity["vs_single_rank"].values()) and parity["all_ranks_same_result"])))
    parity["ok"] = parity_ok

    # ---- secondary measurements (N = 1): the other configs and the steps next to the path ---------------------------
    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = {}
        try:
            del host, hmat
            secondary.update(run_secondary(pkg, torch, h, device, dev, gid))
        except Exception as ex:  # never let a secondary measurement break the contract line
            secondary["error"] = repr(ex)[:300]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r02_pair_kernel_traffic.json")))
        except Exception:
            pass
        value = cmp_dev / (ms_dev * 1e-3)
        e2e = cmp_e2e / (ms_e2e * 1e-3)
        # dominant kernel = the pair kernel (K2); its launches are bracketed by CUDA events on the library's stream
        k2_ms = sum(s["ms_pairs"] for s in st_dev)
        k2_launches = sum(s["pair_launches"] for s in st_dev)
        k2_cmp = float(sum(s["compares"] for s in st_dev))  # rank 0's share
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        n_sms = torch.cuda.get_device_properties(local).multi_processor_count
        B = st_dev[-1]["rank_bits"]
        W = st_dev[-1]["sample_words"]
        # LOP3 lane-ops issued: one per plane the chain runs per 32 sample slots of every evaluated pair, padded slots
        # included -- B + 1 planes, less the empty top plane of the words that skip it (reo_stats.planes_per_word)
        planes = st_dev[-1].get("planes_per_word") or float(B + 1)
        lop3 = k2_cmp / c * (W * 32) * planes / 32.0
        lop3_peak = n_sms * 64 * sm_mhz * 1e6               # the ALU pipe issues 64 LOP3 lanes per clock per SM
        p_cmp = n_sms * 128 * sm_mhz * 1e6                  # SMs x 128 lane-ops/clk x f_SM
        ach_cmp = k2_cmp / (k2_ms * 1e-3) if k2_ms > 0 else 0.0
        ach_lop3 = lop3 / (k2_ms * 1e-3) if k2_ms > 0 else 0.0
        tr = prof.get(args.workload, {})
        roof = {"bound": "alu", "kernel": "reo_pair2_kernel",
                "achieved": ach_lop3 / 1e9, "peak": lop3_peak / 1e9, "unit": "G LOP3 lane-ops/s", "frac": ach_lop3 / lop3_peak,
                "frac_note": "LOP3 lane-ops issued by the borrow chains / (SMs x 64 lanes x SM clock 'of measured' under load); "
                             "the instruction mix (1 POPC + 1 IMAD per B+1 LOP3, 3 LDS.128 per 32) bounds the inner loop alone "
                             "at ~0.86 for 8 planes (scripts/ubench_chain.cu, profiles/r02_ubench_chain.md)",
                "comparisons_per_s": ach_cmp, "p_cmp": p_cmp, "frac_of_p_cmp": ach_cmp / p_cmp,
                "p_cmp_note": "SURVEY 8d compare-ALU roofline: SMs x 128 lane-ops/clk x SM clock, one lane-op per comparison; "
                              "a bit-sliced LOP3 evaluates 32 samples of one plane, so this ratio can exceed 1",
                "traffic": tr.get("dram_bytes_per_launch"), "traffic_note": tr.get("note", "no ncu capture committed for this workload"),
                "launches": k2_launches, "ms_per_launch": k2_ms / max(k2_launches, 1),
                "kernel_share_of_step": k2_ms / ms_dev if ms_dev > 0 else None,
                "rank_bits": B, "sample_words": W, "planes_per_word": planes, "sm_mhz_used": sm_mhz}
        # K1 (rank + bit-plane staging) against the measured HBM copy bandwidth: SURVEY 8d algorithmic bytes =
        # read r*c*8 (Int64 input) + write r*c*(B+1)/8 (bit planes)
        k1_ms = sum(s["ms_stage"] for s in st_dev) / len(st_dev)
        k1_bytes = float(r) * c * 8 + float(r) * c * (B + 1) / 8.0
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roof_k1 = {"bound": "hbm", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None, "peak": hbm_peak,
                   "unit": "GB/s", "frac": (k1_bytes / (k1_ms * 1e-3) / 1e9) / hbm_peak if k1_ms > 0 else None,
                   "traffic": None, "ms": k1_ms, "kernel": "K1 staging (rank + bit planes)",
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs ('of measured')" if "hbm_gbs" in peaks else "fallback 6650 GB/s ('of fallback')",
                   "algorithmic_bytes": k1_bytes}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": args.workload, "note": note, "genes": r, "samples": c,
                       "n_ref_initial": int(ref.sum()), "n_iter": JOB["n_iter"], "n_conv": JOB["n_conv"],
                       "evaluations": st_dev[-1]["iters_done"], "n_deg_per_evaluation": st_dev[-1]["n_deg"],
                       "l2": "flushed between timed steps (256 MB write); the staged planes alone (0.6 GB) exceed L2",
                       "comparisons": "executed (gene pair, sample) evaluations; the symmetric sweep evaluates r(r-1)/2 pairs "
                                      "per all-genes build, as the reference does (src:366-372)",
                       "ordered_triples_per_step": float(sum(s.get("ordered_triples", 0) for s in st_dev)) / len(st_dev) * 1.0,
                       "sharding": f"supertiles of the pair-tile space round-robin over {world} rank(s), tables summed via {args.collective}"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "input": "pinned host memory",
                    "h2d_bytes_per_step": int(nbytes + gid.nbytes + ref.nbytes),
                    "h2d_note": "bytes of the caller's host matrix per step; the library's copy threads narrow integral "
                                "values in 0..65535 to u16 on their way through the pinned bounce buffers, so PCIe carries "
                                "a quarter of an Int64 matrix (csrc/reo_host.cpp; REO_NO_NARROW=1 sends it raw)",
                    "d2h_bytes_per_step": int(r * 15 * 8 + 2 * r), "pageable": e2e_pageable},
            "gpu_launches": int(launches),
            "stage_ms": {"staging": st_dev[-1]["ms_stage"], "pairs": st_dev[-1]["ms_pairs"],
                         "stats": st_dev[-1]["ms_stats"], "total_device": st_dev[-1]["ms_total"],
                         "call_wall": st_dev[-1]["ms_wall"]},
            "ms_per_step_by_rank": ms_by_rank,   # `ms_per_step` is the per-step MAX over ranks
            "last_step_by_rank_pairs_stage_other_total_ms": per_rank,
            "clocks": clocks, "roofline": roof, "roofline_staging": roof_k1, "parity": parity,
        }
        if secondary is not None:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(pkg, co, args.workload)
        print(json.dumps(line))
    h.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit("bench.py: parity check against the oracle FAILED: " + json.dumps(parity))


def run_secondary(pkg, torch, h, device, dev_cells, gid_cells):
    """Short device-side measurements of the other BASELINE configs and of the steps next to the path."""
    L = pkg._lib
    out = {}

    def job(mat, gid, ref, steps=5, warmup=2):
        ms, st = [], None
        for i in range(warmup + steps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = h.identify_degs(mat, gid, 2, ref, JOB["pval_reo"], JOB["pval_deg"], JOB["padj_deg"], JOB["n_iter"], JOB["n_conv"])
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ms.append(e0.elapsed_time(e1)); st = o.stats
        t = sum(ms) / len(ms)
        return {"ms_per_job": t, "value": st["compares"] / (t * 1e-3), "unit": UNIT, "evaluations": st["iters_done"],
                "pairs_ms": st["ms_pairs"], "staging_ms": st["ms_stage"], "stats_ms": st["ms_stats"], "rank_bits": st["rank_bits"],
                "lop3_pipe_frac": (st["compares"] / mat_c(mat) * st["sample_words"] * (st.get("planes_per_word") or (st["rank_bits"] + 1))) / (st["ms_pairs"] * 1e-3)
                                  / (torch.cuda.get_device_properties(0).multi_processor_count * 64 * 1.965e9)
                if st["ms_pairs"] > 0 and st["rank_bits"] > 0 else None}

    def mat_c(mat):
        return mat.c if isinstance(mat, pkg.DeviceMatrix) else mat.shape[1]

    def on_device(data):
        t = torch.from_numpy(np.ascontiguousarray(data.T)).to(device)
        return pkg.DeviceMatrix(t.data_ptr(), L.REO_I64 if data.dtype == np.int64 else L.REO_F64, data.shape[0], data.shape[1],
                                data.shape[0], keepalive=t)

    # configs[1]: bulk 20k x (100 + 100), 3000 HK references -- the small, latency-dominated job
    data, _, gid, ref = make_workload(pkg, "c2_bulk_20kx200")
    out["small_job_c2_bulk_20kx200"] = job(on_device(data), gid, ref, steps=10, warmup=3)
    # float path: the same bulk matrix as TPM-like non-integral values -> raw-FP64 compare kernel
    tpm = data.astype(np.float64) / np.maximum(data.sum(axis=0, keepdims=True), 1) * 1e6 + 0.013
    out["float_path_tpm_20kx200"] = job(on_device(np.asfortranarray(tpm)), gid, ref, steps=3, warmup=1)
    del data, tpm
    # configs[0]: the reference's bundled test data (19999 genes x 5 vs 5), seeded 3000-gene reference mask
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "bundled_c1.npz"))
        d1 = np.asfortranarray(z["data"].astype(np.int64))
        out["c1_bundled_19999x10"] = job(on_device(d1), z["gid"].astype(np.int32), z["ref"].astype(bool), steps=10, warmup=3)
    except Exception as ex:
        out["c1_bundled_19999x10"] = {"skipped": repr(ex)[:120]}
    # configs[2] from cells: 25k genes x (10k + 10k) cells -> reo_pseudobulk (n_pseudo = 50 per group) ->
    # identify_degs on the 25k x 100 pseudo-bulk matrix without leaving HBM
    r3, c3 = 25000, 20000
    cells = dev_cells[:, :r3].contiguous()                  # [c, r3] of the resident single-cell matrix
    dm = pkg.DeviceMatrix(cells.data_ptr(), L.REO_I64, r3, c3, r3, keepalive=cells)
    rng = np.random.default_rng(1234)
    profiles = []
    for g in (0, 1):
        idx = rng.permutation(np.nonzero(np.asarray(gid_cells) == g)[0])
        profiles += [idx[i::50] for i in range(50)]         # 50 pseudo-bulk profiles of 200 cells per group
    gid_pb = np.repeat(np.arange(2, dtype=np.int32), 50)
    ms = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, pb = h.pseudobulk(dm, profiles, to_host=False)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
    t_pb = min(ms[1:])
    ref3 = np.zeros(r3, bool)
    ref3[np.random.default_rng(4321).choice(r3, 3000, replace=False)] = True
    j3 = job(pb, gid_pb, ref3, steps=5, warmup=2)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    gbs = r3 * c3 * 8 / (t_pb * 1e-3) / 1e9
    out["c3_pseudobulk_from_cells_25kx20k"] = {
        "pseudobulk_ms": t_pb, "pseudobulk_GBps": gbs, "hbm_frac": gbs / float(peaks.get("hbm_gbs", 6650.0)),
        "note": "reo_pseudobulk call wall time incl. upload of the cell lists; algorithmic bytes = one read of the 4.0 GB matrix",
        "identify_degs_on_pseudobulk": j3}
    return out


if __name__ == "__main__":
    main()
