#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  python scripts/profile_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python scripts/profile_summary.py kernel   gpurun_out/prof.ncu-rep   profiles/r01_pair_kernel.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0][:70]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1.0)
        agg[k][0] += 1
        agg[k][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write(f"Source: `{src}`; {n} launches, {tot / 1e6:.3f} ms of kernel time. Per-launch times under ncu are "
                "cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / tot:.3f} |\n")


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
]


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(raw.splitlines()))
    hdr, units = rd[0], rd[1]
    stall = [k for k in hdr if "smsp__average_warps_issue_stalled" in k and "per_issue_active" in k]
    srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                          text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(srcp.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = []
            blocks.append((row[1] if len(row) > 1 else "?", cur))
            continue
        if cur is not None:
            cur.append(row)
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture: `{src}`\n\n")
        for li, row in enumerate(rd[2:]):
            d = dict(zip(hdr, row))
            f.write(f"## launch {li}: `{d.get('Kernel Name', '?')}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS + stall:
                if k in d and d[k] not in ("",):
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            if li < len(blocks):
                name, blk = blocks[li]
                h2, rows = blk[0], blk[1:]
                iS, iE, iM = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
                agg, samp, tot, tots = collections.Counter(), collections.Counter(), 0, 0
                for r in rows:
                    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
                    op = m.group(2) if m else "?"
                    agg[op] += int(r[iE]); samp[op] += int(r[iM]); tot += int(r[iE]); tots += int(r[iM])
                f.write(f"\nSASS opcode mix ({tot} warp-instructions, {tots} stall samples):\n\n"
                        "| opcode | executed | share | stall-sample share |\n|---|---:|---:|---:|\n")
                for op, n in agg.most_common(14):
                    f.write(f"| {op} | {n} | {n / max(tot, 1):.3f} | {samp[op] / max(tots, 1):.3f} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
