#!/usr/bin/env python
"""
Does the product's tie-coin rule change DEG calls?  

The reference flips an independent rand(Bool) per (unordered gene pair, sample) on a tie (src/RankCompV3.jl:72-73).  The
product's rule  coin(i,j,s) = u(i,s) ^ u(j,s) ^ [i<j]  is a fair coin for every pair and is mirrored exactly, but all
coins of one sample come from G bits: coin(i,j)^coin(i,k)^coin(j,k) is constant.  Per-gene tables are equal in
distribution; the JOINT distribution across genes, which the pooled empirical-null std (src:409-411) sees, is not.
This script measures the effect with the CPU oracle (test infrastructure): identify_degs under both rules
(oracle coin mode 0 = XOR rule, 1 = independent per-pair hash coins) over N seeds, on the reference's bundled test
data and on a zero-inflated single-cell-like 2000 x 400 matrix, and reports
  * Jaccard index of the DEG sets between the two RULES at the same seed,
  * Jaccard index between two SEEDS under the same rule (the seed-to-seed noise either rule has anyway),
  * the empirical-null se of the final evaluation and the DEG counts under each rule.
Usage: python scripts/coin_study.py [n_seeds] > profiles/r02_coin_study.md
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
_, co = ge.load_oracle()
co.use_all_cores()
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20


def jaccard(a, b):
    u = np.logical_or(a, b).sum()
    return float(np.logical_and(a, b).sum() / u) if u else 1.0


def run(data, gid, ref, seed, mode):
    co.set_coin_mode(mode)
    try:
        thr = co.thresholds_for(gid, 2, 0.01)
        out = co.identify_degs(data, gid, 2, thr, 1.0, 0.05, ref, 128, 5, seed=seed)
    finally:
        co.set_coin_mode(0)
    res = out["result"][0]
    se, _ = co.empirical_null(res[:, 11])
    return out["updown"][0] != 0, out["updown"][0], se, out["iters"][0]


def study(name, data, gid, ref):
    calls = {0: [], 1: []}
    signs = {0: [], 1: []}
    ses = {0: [], 1: []}
    for s in range(n_seeds):
        for mode in (0, 1):
            deg, ud, se, it = run(data, gid, ref, 1000 + s, mode)
            calls[mode].append(deg); signs[mode].append(ud); ses[mode].append(se)
    between_rules = [jaccard(calls[0][s], calls[1][s]) for s in range(n_seeds)]
    within = {m: [jaccard(calls[m][s], calls[m][(s + 1) % n_seeds]) for s in range(n_seeds)] for m in (0, 1)}
    flips = [int(np.sum((signs[0][s] != 0) & (signs[1][s] != 0) & (signs[0][s] != signs[1][s]))) for s in range(n_seeds)]
    ms = lambda v: f"{np.mean(v):.4f} +- {np.std(v):.4f}"  # noqa: E731
    ties = float(np.mean(data[:, None, :8] == data[None, :200, :8])) if data.shape[0] <= 4000 else float("nan")
    print(f"### {name}: {data.shape[0]} genes x {data.shape[1]} samples, {int(ref.sum())} initial reference genes, {n_seeds} seeds\n")
    print("| quantity | XOR rule (product) | independent per-pair coins |")
    print("|---|---|---|")
    print(f"| DEGs called | {ms([c.sum() for c in calls[0]])} | {ms([c.sum() for c in calls[1]])} |")
    print(f"| empirical-null se (final evaluation) | {ms(ses[0])} | {ms(ses[1])} |")
    print(f"| Jaccard of DEG sets, seed s vs seed s+1, same rule | {ms(within[0])} | {ms(within[1])} |")
    print(f"| Jaccard of DEG sets, XOR rule vs per-pair coins at the same seed | {ms(between_rules)} | |")
    print(f"| genes called in both with opposite direction | {sum(flips)} in {n_seeds} seeds | |")
    if ties == ties:
        print(f"\nFraction of tied (pair, sample) comparisons in a 200-gene x 8-sample probe: {ties:.3f}.")
    print()
    return between_rules, within


print("# Tie-coin rule: XOR of per-gene bits vs independent per-pair coins (CPU oracle, scripts/coin_study.py)\n")
z = np.load(os.path.join(ROOT, "tests", "golden", "bundled_c1.npz"))
study("bundled test data (test/fn_expr.txt)", z["data"].astype(np.int64), z["gid"].astype(np.int32), z["ref"].astype(bool))
data, group, is_de = pkg.synth.scrna(2000, 200, 200, seed=77)
_, gid = pkg.api.group_levels(group)
ref = pkg.synth.random_mask(2000, 300, seed=4321)
study("zero-inflated single-cell-like counts", data, gid, ref)
