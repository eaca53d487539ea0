#!/usr/bin/env python
"""
bench.py -- the REO hot path (identify_degs, src/RankCompV3.jl:339-438) on BASELINE.json's metric:
gene-pair x sample comparisons per second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libreo_cuda.so)
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference (oracle/)

A "step" is one complete identify_degs job over one synthetic expression matrix.  Default workload = BASELINE.json
configs[1]: bulk RNA-seq, 20 000 genes x (100 vs 100) samples, 3 000 house-keeping reference genes, n_iter = 128.
  value : W_ord / t with the matrix already resident in HBM (raw Int64, Julia layout) -- staging, every pair-kernel
          launch, every statistics kernel and the result read-back are inside the timed region.
  e2e   : the same job through the public call with HOST (pinned) buffers: H2D of the matrix inside the timed region.
W_ord = ordered (gene, reference gene, sample) triples whose REO was evaluated, as counted by the library.
N > 1 (torchrun, one process per GPU): gene-row tiles are sharded across ranks, tables all-gathered over NCCL once
per evaluation -- the total work is fixed, i.e. strong scaling.
The Julia reference cannot run in this image (no julia binary), so --impl reference and cpu_baseline time the plain-C
restatement oracle/reo_oracle.c ("port") on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
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
    "c2_bulk_20kx200": ("bulk", 20000, 100, 100, 3000, "BASELINE configs[1]: bulk 20k genes x 100 vs 100, 3000 HK refs"),
    "c3_pseudobulk_25kx100": ("bulk", 25000, 50, 50, 3000, "BASELINE configs[2] core shape: 25k genes x 50 vs 50"),
    "c4_scrna_30kx20k": ("scrna", 30000, 10000, 10000, 3000, "BASELINE configs[3]: 30k genes x 10k vs 10k cells"),
    "c5_allref_30kx20k": ("scrna", 30000, 10000, 10000, 0, "BASELINE configs[4]: all genes as references"),
    "mid_scrna_8kx4k": ("scrna", 8000, 2000, 2000, 1500, "mid-size single-cell: 8k genes x 2k vs 2k cells"),
    "tiny": ("bulk", 2000, 20, 20, 300, "smoke-sized"),
}


def make_workload(pkg, name, device=None):
    """Returns (column-major r x c Int64 numpy matrix, group ids, reference mask).  Single-cell shapes are
    generated on the GPU when one is given (same model, seconds instead of minutes) and copied back."""
    kind, r, n1, n2, n_ref, _ = WORKLOADS[name]
    if kind == "bulk":
        data, group, is_de = pkg.synth.bulk(r, n1, n2)
    elif device is not None:
        t, group, is_de = pkg.synth.scrna_torch(r, n1, n2, device=device)
        data = t.cpu().numpy().T          # [c, r] row-major -> F-ordered r x c view
        del t
    else:
        data, group, is_de = pkg.synth.scrna(r, n1, n2)
    ref = pkg.synth.reference_mask(is_de, n_ref) if n_ref > 0 else np.ones(r, dtype=bool)
    levels, gid = pkg.api.group_levels(group)
    return np.asfortranarray(data), gid, ref


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # samples under load only: the job is short, idle samples sit at the base clock
        hi = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(hi) if hi else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(pkg, co, data, gid, ref, target_s=12.0):
    """Time the plain-C restatement on a bounded sample: a block of gene rows against the workload's reference
    columns (iteration-0 table build of the same matrix), all host threads."""
    thr = co.thresholds_for(gid, 2, 0.01)
    cols = np.nonzero(ref)[0]
    r, c = data.shape
    d = np.asfortranarray(data.astype(np.float64))
    rows = 64
    t0 = time.perf_counter()
    _, n = co.block_tables(d, gid, 2, thr, cols, seed=pkg.synth.TIE_SEED, i0=0, i1=rows)
    dt = time.perf_counter() - t0
    rate = n / dt
    rows = int(max(64, min(r, rows * target_s / max(dt, 1e-3) * 0.8)))
    t0 = time.perf_counter()
    _, n = co.block_tables(d, gid, 2, thr, cols, seed=pkg.synth.TIE_SEED, i0=0, i1=rows)
    dt = time.perf_counter() - t0
    return dict(value=n / dt, unit=UNIT, cores=co.num_threads(), kind="port",
                sample=f"{rows} gene rows x {len(cols)} reference genes x {c} samples = {n:.3e} comparisons in {dt:.2f} s "
                       f"(oracle/reo_oracle.c, OpenMP; the Julia reference cannot run here)"), rows


def run_reference(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, co = ge.load_oracle()
    co.use_all_cores()                      # torchrun exports OMP_NUM_THREADS=1
    data, gid, ref = make_workload(pkg, args.workload)
    thr = co.thresholds_for(gid, 2, 0.01)
    cols = np.nonzero(ref)[0]
    d = np.asfortranarray(data.astype(np.float64))
    base, rows = cpu_baseline(pkg, co, data, gid, ref, target_s=max(2.0, 60.0 / max(args.steps + args.warmup, 1)))
    times, n = [], 0
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, n = co.block_tables(d, gid, 2, thr, cols, seed=pkg.synth.TIE_SEED, i0=0, i1=rows)
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    val = n / t
    base["value"] = val
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "note": WORKLOADS[args.workload][5],
                   "sample": f"each step = {rows} gene rows x {len(cols)} refs x {data.shape[1]} samples"},
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_bulk_20kx200", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probe", action="store_true",
                    help="skip the secondary device-only measurement on BASELINE configs[4] (30k x 30k x 20k, all genes as "
                         "references), reported under 'scaling_probe'")
    ap.add_argument("--collective", default="nccl", choices=["nccl", "torch"],
                    help="N>1: all-gather by NCCL inside the library (default) or through torch.distributed")
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)

    data, gid, ref = make_workload(pkg, args.workload, device=f"cuda:{local}")
    r, c = data.shape
    h = pkg.Reo(local, seed=pkg.synth.TIE_SEED)
    if world > 1:
        from importlib import import_module
        dmod = import_module(pkg.__name__ + ".dist")
        if args.collective == "nccl":
            dmod.init_nccl_in_library(h, rank, world)
        else:
            h.set_collective(rank, world, dmod.make_torch_allgather(rank, world))

    # inputs: pinned host copy (e2e) and an HBM-resident copy (value); column-major r x c Int64 = Julia's Matrix
    host = torch.from_numpy(np.ascontiguousarray(data.T)).pin_memory()     # [c, r] row-major == r x c column-major
    dev = host.to(f"cuda:{local}")
    dmat = pkg.DeviceMatrix(dev.data_ptr(), pkg._lib.REO_I64, r, c, r, keepalive=dev)
    hmat = host.numpy().T                                                    # F-ordered view of the pinned buffer
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def one(mat):
        return h.identify_degs(mat, gid, 2, ref, 0.01, 1.0, 0.05, 128, 5)

    def timed(mat, steps, warmup):
        out = None
        for _ in range(warmup):
            one(mat)
        total_ms, stats = 0.0, []
        by_rank = [0.0] * world
        for _ in range(steps):
            flush.fill_(1)                      # L2 flush between timed iterations (inputs are 32 MB < L2)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            out = None                          # release the previous result buffers before the next call
            e0.record()
            out = one(mat)                      # blocking: returns after the result read-back
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
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
        timed.by_rank = by_rank
        return total_ms, stats, out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, st_dev, out_dev = timed(dmat, args.steps, args.warmup)
    ms_by_rank = [round(x, 4) for x in timed.by_rank]
    if os.environ.get("REO_BENCH_RANKLOG"):      # diagnosis: every rank's own view of its last timed step
        with open(os.path.join(os.environ["REO_BENCH_RANKLOG"], f"rank_{rank}.json"), "w") as f:
            json.dump({"rank": rank, "ms_by_rank": ms_by_rank, "last_step": st_dev[-1]}, f)
    ms_e2e, st_e2e, out_e2e = timed(hmat, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # whole-job compares: every rank evaluated its own row shard
    cmp_dev = torch.tensor([float(sum(s["compares"] for s in st_dev))], device=f"cuda:{local}", dtype=torch.float64)
    cmp_e2e = torch.tensor([float(sum(s["compares"] for s in st_e2e))], device=f"cuda:{local}", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(cmp_dev)
        dist.all_reduce(cmp_e2e)
    assert np.array_equal(out_dev.result[:, :, 2:11], out_e2e.result[:, :, 2:11])

    # secondary, device-only measurement on the shape BASELINE.json names for the 1/2/4/8-GPU sweep (configs[4]):
    # the primary workload is a ~4 ms job whose replicated part bounds strong scaling; this one shows the pair kernel
    # under row-tile sharding on 1.8e13 comparisons per table build.  Not part of `value`.
    probe = None
    if not args.no_probe and args.workload == "c2_bulk_20kx200":
        try:
            del dev, dmat, host, hmat
            t, group5, _ = pkg.synth.scrna_torch(30000, 10000, 10000, device=f"cuda:{local}")
            _, gid5 = pkg.api.group_levels(group5)
            ref5 = np.ones(30000, dtype=bool)
            dm5 = pkg.DeviceMatrix(t.data_ptr(), pkg._lib.REO_I64, 30000, 20000, 30000, keepalive=t)
            ms5, cmp5, st5 = 0.0, 0.0, None
            for it in range(2 + 3):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                o5 = h.identify_degs(dm5, gid5, 2, ref5, 0.01, 1.0, 0.05, 128, 5)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    m = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
                    if dist is not None:
                        dist.all_reduce(m, op=dist.ReduceOp.MAX)
                    ms5 += float(m.item())
                    cmp5 += o5.stats["compares"]
                    st5 = o5.stats
                o5 = None
            c5 = torch.tensor([cmp5], device=f"cuda:{local}", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(c5)
            probe = {"workload": "c5_allref_30kx20k", "note": WORKLOADS["c5_allref_30kx20k"][5], "steps": 3, "warmup": 2,
                     "value": c5.item() / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5 / 3,
                     "inputs": "resident in HBM (1.2 GB staged > L2)", "rank_bits": st5["rank_bits"],
                     "pairs_ms": st5["ms_pairs"], "staging_ms": st5["ms_stage"], "evaluations": st5["iters_done"]}
            del t, dm5
        except Exception as ex:  # never let the probe break the contract line
            probe = {"workload": "c5_allref_30kx20k", "error": repr(ex)[:200]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        value = cmp_dev.item() / (ms_dev * 1e-3)
        e2e = cmp_e2e.item() / (ms_e2e * 1e-3)
        # dominant kernel = the pair kernel (K2); its launches are bracketed by CUDA events on the library's stream
        k2_ms = sum(s["ms_pairs"] for s in st_dev)
        k2_launches = sum(s["pair_launches"] for s in st_dev)
        k2_cmp = float(sum(s["compares"] for s in st_dev))  # rank 0's shard
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        n_sms = torch.cuda.get_device_properties(local).multi_processor_count
        p_cmp = n_sms * 128 * sm_mhz * 1e6                  # P_cmp = SMs x 128 lane-ops/clk x f_SM
        ach = k2_cmp / (k2_ms * 1e-3) if k2_ms > 0 else 0.0
        B = st_dev[-1]["rank_bits"]
        W = st_dev[-1]["sample_words"]
        # LOP3 lane-ops actually issued: (B+1) per 32 sample slots, padded words included
        lop3 = k2_cmp / c * (W * 32) * (B + 1) / 32.0
        # DRAM traffic of the dominant launch from the ncu --set full capture in profiles/r01_pair_kernel_ncu.md
        # (dram__bytes_read.sum + dram__bytes_write.sum, delta build of the default workload: 13.6 MB ~= the staged
        # row planes + column panel read once; everything else is served by L2, hit rate 98 %)
        traffic = 13.6e6 if args.workload == "c2_bulk_20kx200" else None
        roof = {"bound": "alu", "achieved": ach / 1e9, "peak": p_cmp / 1e9, "unit": "Gcmp/s", "frac": ach / p_cmp,
                "traffic": traffic, "traffic_note": "bytes per launch of the largest pair-kernel launch (ncu, profiles/)",
                "note": "compare-ALU roofline of BASELINE.md (SMs x 128 lanes x measured SM clock, 'of measured' clock; "
                        "fallback 1965 MHz if nvidia-smi gave no sample); bit-sliced LOP3 evaluates 32 samples per lane-op",
                "kernel": "reo_pair_kernel", "launches": k2_launches, "ms_per_launch": k2_ms / max(k2_launches, 1),
                "kernel_share_of_step": k2_ms / ms_dev if ms_dev > 0 else None,
                "lop3_pipe_frac": (lop3 / (k2_ms * 1e-3)) / (n_sms * 64 * sm_mhz * 1e6) if k2_ms > 0 else None,
                "rank_bits": B, "sample_words": W, "sm_mhz_used": sm_mhz}
        # K1 (rank + bit-plane staging) against the measured HBM copy bandwidth: SURVEY 8d algorithmic bytes =
        # read r*c*8 (Int64 input) + write r*c*2 (dense ranks).  Not the dominant kernel (3-4 % of a job).
        k1_ms = sum(s["ms_stage"] for s in st_dev) / len(st_dev)
        k1_bytes = float(r) * c * 8 + float(r) * c * 2
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roof_k1 = {"bound": "hbm", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None, "peak": hbm_peak,
                   "unit": "GB/s", "frac": (k1_bytes / (k1_ms * 1e-3) / 1e9) / hbm_peak if k1_ms > 0 else None,
                   "traffic": None, "ms": k1_ms, "kernel": "rank_columns_kernel + bitplanes_kernel",
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                   "note": "latency-bound at this size: one 1024-thread CTA per sample column (200 columns = 1.35 waves of "
                           "148 SMs with the 185 KB presence bitmap bulk counts need); single-cell sized inputs take the "
                           "small-bitmap tier (2 CTAs/SM, column read from DRAM once) and reach ~1.1 TB/s of Int64 input"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": args.workload, "note": WORKLOADS[args.workload][5], "genes": r, "samples": c,
                       "n_ref_initial": int(ref.sum()), "n_iter": 128, "n_conv": 5,
                       "evaluations": st_dev[-1]["iters_done"], "n_deg_per_evaluation": st_dev[-1]["n_deg"],
                       "l2": "flushed between timed steps (256 MB write); inputs are 32 MB",
                       "sharding": f"gene-row tiles over {world} rank(s), all-gather of per-gene tables via {args.collective}"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(data.nbytes + gid.nbytes + ref.nbytes),
                    "d2h_bytes_per_step": int(r * 15 * 8 + 2 * r)},
            "gpu_launches": int(sum(s["kernel_launches"] for s in st_dev)),
            "stage_ms": {"staging": st_dev[-1]["ms_stage"], "pairs": st_dev[-1]["ms_pairs"],
                         "stats": st_dev[-1]["ms_stats"], "total_device": st_dev[-1]["ms_total"],
                         "call_wall": st_dev[-1]["ms_wall"]},
            "ms_per_step_by_rank": ms_by_rank,   # `ms_per_step` is the per-step MAX over ranks
            "clocks": clocks, "roofline": roof, "roofline_staging": roof_k1,
        }
        if probe is not None:
            line["scaling_probe"] = probe
        if not args.no_cpu_baseline:
            _, co = ge.load_oracle()
            co.use_all_cores()
            line["cpu_baseline"], _ = cpu_baseline(pkg, co, data, gid, ref)
        print(json.dumps(line))
    h.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
