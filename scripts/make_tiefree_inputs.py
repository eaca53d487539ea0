#!/usr/bin/env python
"""Writes the tie-free inputs julia/dump_reference_fixture.jl feeds to the UNMODIFIED reference
(tests/golden/julia_in/<case>_{expr,meta,ref,par}.tsv).  Tie-free twins (synth.tie_free: every sample column is a
permutation of distinct integers with the DE structure of the synthetic model) never reach rand(Bool) in is_greater
(src/RankCompV3.jl:72-73), so the reference's output on them is deterministic.
    python scripts/make_tiefree_inputs.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
OUT = os.path.join(ROOT, "tests", "golden", "julia_in")
os.makedirs(OUT, exist_ok=True)

CASES = {
    # name: (kind, genes, group sizes, reference genes, (pval_reo, pval_deg, padj_deg, n_iter, n_conv))
    "bulk_600x26": ("bulk", 600, (12, 14), 120, (0.01, 1.0, 0.05, 128, 5)),
    "scrna_500x80": ("scrna", 500, (40, 40), 100, (0.01, 1.0, 0.05, 128, 5)),
    "bulk3_300x36": ("bulk", 300, (10, 12, 14), 60, (0.01, 1.0, 0.05, 128, 5)),
    "bulk_allref_400x20": ("bulk", 400, (10, 10), 400, (0.01, 1.0, 0.05, 3, 0)),
}


def build(name):
    kind, r, sizes, n_ref, par = CASES[name]
    if kind == "bulk":
        data, group, is_de = pkg.synth.bulk(r, sizes[0], sum(sizes[1:]), seed=4242)
    else:
        data, group, is_de = pkg.synth.scrna(r, sizes[0], sum(sizes[1:]), seed=4242)
    data = pkg.synth.tie_free(data, seed=4242)
    group = sum([[f"g{k + 1}"] * n for k, n in enumerate(sizes)], [])
    ref = pkg.synth.random_mask(r, n_ref, seed=4321) if n_ref < r else np.ones(r, bool)
    return data, group, ref, par


def main():
    for name in CASES:
        data, group, ref, par = build(name)
        r, c = data.shape
        genes = [f"G{i + 1:05d}" for i in range(r)]
        samples = [f"S{s + 1:04d}" for s in range(c)]
        with open(os.path.join(OUT, f"{name}_expr.tsv"), "w") as f:
            f.write("gene\t" + "\t".join(samples) + "\n")
            for i in range(r):
                f.write(genes[i] + "\t" + "\t".join(str(int(v)) for v in data[i]) + "\n")
        with open(os.path.join(OUT, f"{name}_meta.tsv"), "w") as f:
            f.write("Name\tGroup\n")
            for s in range(c):
                f.write(f"{samples[s]}\t{group[s]}\n")
        with open(os.path.join(OUT, f"{name}_ref.tsv"), "w") as f:
            f.write("gene\tis_ref\n")
            for i in range(r):
                f.write(f"{genes[i]}\t{int(ref[i])}\n")
        with open(os.path.join(OUT, f"{name}_par.tsv"), "w") as f:
            f.write("pval_reo\tpval_deg\tpadj_deg\tn_iter\tn_conv\n")
            f.write("\t".join(repr(v) for v in par) + "\n")
        print(name, data.shape, "refs", int(ref.sum()))


if __name__ == "__main__":
    main()
