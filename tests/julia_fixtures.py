"""Loader for the reference-run fixtures: inputs under tests/golden/julia_in (scripts/make_tiefree_inputs.py) and, when a
maintainer with Julia has produced them, the UNMODIFIED reference's outputs under tests/golden/julia_out
(julia/dump_reference_fixture.jl).  The image this repo is built in has no Julia, so the outputs may be missing."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
IN_DIR = os.path.join(HERE, "golden", "julia_in")
OUT_DIR = os.path.join(HERE, "golden", "julia_out")
CODES = {"up": 1, "down": -1, "no change": 0}


def cases():
    return sorted(f[:-len("_expr.tsv")] for f in os.listdir(IN_DIR) if f.endswith("_expr.tsv"))


def load_input(case):
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(IN_DIR, case + "_expr.tsv"))]
    genes = [r[0] for r in rows[1:]]
    data = np.array([[int(v) for v in r[1:]] for r in rows[1:]], dtype=np.int64)
    group = [l.rstrip("\n").split("\t")[1] for l in open(os.path.join(IN_DIR, case + "_meta.tsv"))][1:]
    ref = np.array([int(l.rstrip("\n").split("\t")[1]) for l in open(os.path.join(IN_DIR, case + "_ref.tsv")).readlines()[1:]],
                   dtype=bool)
    par = open(os.path.join(IN_DIR, case + "_par.tsv")).readlines()[1].split("\t")
    par = (float(par[0]), float(par[1]), float(par[2]), int(par[3]), int(par[4]))
    return genes, data, group, ref, par


def load_reference_output(case):
    """-> (result [K, r, 15] float64, updown [K, r] int8) as dumped from the reference's `res` matrix, or None if absent."""
    path = os.path.join(OUT_DIR, case + ".tsv")
    if not os.path.exists(path):
        return None
    rows = [l.rstrip("\n").split("\t") for l in open(path) if l.strip()]
    ncol = len(rows[0]) - 1
    assert ncol % 16 == 0, "expected 1 + 16K columns (src:394/430)"
    K, r = ncol // 16, len(rows)
    result = np.zeros((K, r, 15))
    updown = np.zeros((K, r), dtype=np.int8)
    for i, row in enumerate(rows):
        for k in range(K):
            result[k, i] = [float(v) for v in row[1 + 16 * k:16 + 16 * k]]
            updown[k, i] = CODES[row[16 + 16 * k]]
    return result, updown


def compare(result, updown, want_result, want_updown, rtol=1e-12):
    """The bars of north_star: tables bit-exact, p-values within 1e-12 relative, identical calls."""
    assert result.shape == want_result.shape, (result.shape, want_result.shape)
    assert np.array_equal(result[:, :, 2:11], want_result[:, :, 2:11]), "contingency tables differ from the reference"
    assert np.array_equal(updown, want_updown), "up/down/no-change calls differ from the reference"
    for col in (0, 1):
        a, b = result[:, :, col], want_result[:, :, col]
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        assert rel.max() <= rtol, (col, rel.max())
    for col in (11, 12, 13, 14):
        a, b = result[:, :, col], want_result[:, :, col]
        assert np.all(np.abs(a - b) <= rtol * np.abs(b) + 4e-15), col
