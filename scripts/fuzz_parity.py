#!/usr/bin/env python
"""Randomised parity sweep (GPU): many shapes / group layouts / dtypes / masks against the C oracle.
    python scripts/fuzz_parity.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
oracle, co = ge.load_oracle()
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = pkg.Reo(0, seed=pkg.synth.TIE_SEED)
t0 = time.time()
bad = 0
for case in range(n_cases):
    r = int(rng.integers(11, 700)) if rng.random() < 0.9 else int(rng.integers(700, 3500))
    gnum = int(rng.choice([2, 2, 2, 3, 4]))
    sizes = [int(rng.choice([1, 2, 5, 31, 32, 33, 64, 65, 100, int(rng.integers(1, 200))])) for _ in range(gnum)]
    if rng.random() < 0.1:
        sizes[0] = int(rng.integers(600, 1400))      # compare path (no lookup tables)
        r = min(r, 150)
    c = sum(sizes)
    kind = rng.choice(["counts", "sparse", "wide", "float", "tiefree", "const"])
    if kind == "counts":
        data = rng.poisson(np.exp(rng.normal(2, 1.5, r))[:, None] * np.ones((1, c))).astype(np.int64)
    elif kind == "sparse":
        data = (rng.poisson(0.3, size=(r, c)) * (rng.random((r, c)) < 0.5)).astype(np.int64)
    elif kind == "wide":
        data = rng.integers(-2**40, 2**40, size=(r, c)).astype(np.int64)
        data[:, ::2] //= 2**30
    elif kind == "float":
        data = np.round(rng.normal(3, 1, size=(r, c)), 2)
    elif kind == "tiefree":
        data = np.stack([rng.permutation(r) for _ in range(c)], axis=1).astype(np.int64)
    else:
        data = np.full((r, c), 3, dtype=np.int64)
    dtype = rng.choice(["i64", "i32", "f64"]) if kind not in ("float", "wide") else ("f64" if kind == "float" else "i64")
    arr = data.astype({"i64": np.int64, "i32": np.int32, "f64": np.float64}[dtype]) if kind != "float" else data
    labels = np.repeat(np.arange(gnum), sizes)
    perm = rng.permutation(c)
    arr = arr[:, perm]
    group = [f"g{labels[p]}" for p in perm]
    levels, gid = oracle.group_levels(group)
    thr = co.thresholds_for(gid, gnum, 0.01)
    frac = rng.choice([0.0, 0.05, 0.3, 0.9, 1.0])
    mask = rng.random(r) < frac
    ref64 = np.asarray(arr, dtype=np.float64)
    try:
        want = co.identify_degs(ref64, gid, gnum, thr, 1.0, 0.05, mask, 6, 2, seed=7)
        out = h.identify_degs(arr, gid, gnum, mask, 0.01, 1.0, 0.05, 6, 2)
        ok = (out.iters == want["iters"] and np.array_equal(out.result[:, :, 2:11], want["result"][:, :, 2:11])
              and np.array_equal(out.updown, want["updown"]) and np.array_equal(out.final_ref, want["final_ref"]))
        a, b = out.result[:, :, :2], want["result"][:, :, :2]
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        rel[a == b] = 0
        ok = ok and np.nanmax(rel) <= 1e-12
        # d1 can cancel to ~1e-16 (one ulp of a log), and z1 = d1 / se magnifies that by 1/se
        atol = np.array([4e-15, 4e-15, 4e-15, 1e-13])
        st = np.abs(out.result[:, :, 11:15] - want["result"][:, :, 11:15]) <= 1e-12 * np.abs(want["result"][:, :, 11:15]) + atol
        ok = ok and bool(st.all())
    except Exception as ex:  # noqa: BLE001
        ok = False
        print("EXC", repr(ex)[:200])
    if not ok:
        bad += 1
        print(f"MISMATCH case {case}: r={r} sizes={sizes} kind={kind} dtype={dtype} frac={frac}")
        try:
            print("   iters", out.iters, want["iters"], "tables equal", np.array_equal(out.result[:, :, 2:11], want["result"][:, :, 2:11]),
                  "updown equal", np.array_equal(out.updown, want["updown"]), "ref equal", np.array_equal(out.final_ref, want["final_ref"]))
            for k in range(out.result.shape[0]):
                d = np.argwhere(~np.isclose(out.result[k], want["result"][k], rtol=1e-12, atol=4e-15, equal_nan=True))
                print(f"   level {k}: {len(d)} cells differ; first", [(int(i), int(j), float(out.result[k, i, j]), float(want["result"][k, i, j])) for i, j in d[:4]],
                      "n_deg", out.stats["n_deg"], "se", float(out.result[k, 0, 13]))
        except Exception as ex:  # noqa: BLE001
            print("   detail failed", repr(ex)[:200])
print(f"{n_cases} cases, {bad} mismatches, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
