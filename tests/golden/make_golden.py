"""
Generates the committed fixtures under tests/golden/ (run in the build container, where /root/reference
exists):  python tests/golden/make_golden.py

 kat_mccullagh.json    the reference's only known-answer vector (src/RankCompV3.jl:206-222,
                       test/McCullagh_test.jl:38-39), copied as data.
 small_case.npz        a seeded 400 x (13+19) count matrix with ties + the oracle's full identify_degs output.
 bundled_c1.npz        the reference's bundled test input (test/fn_expr.txt, test/fn_meta.txt; 19999 genes after
                       the all-zero row is dropped, 5 vs 5) with a seeded 3000-gene reference mask, and the
                       C oracle's output for it (tables, p-values, calls).  Tie seed 7, mask seed 4321.
Oracle outputs here are produced by oracle/reo_oracle.c; oracle/reo_oracle.py is checked against them in
tests/test_oracle.py ("two independent restatements agree").
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from conftest import small_case  # noqa: E402

oracle, co = ge.load_oracle()
pkg = ge.load_package()

# 1. KAT
kat = dict(mat=[[43, 8, 3, 0], [2, 2, 5, 3], [1, 0, 7, 2], [0, 0, 1, 5]],
           expected=[0.005469174895116946, 1.4504988072997458, 1.502600073417028, 0.5221345956920705,
                     2.778017046308073],
           N=[[14, 4, 0], [4, 12, 3], [0, 3, 6]], R=[11, 11, 5],
           source="src/RankCompV3.jl:206-222; test/McCullagh_test.jl:38-39")
json.dump(kat, open(os.path.join(HERE, "kat_mccullagh.json"), "w"), indent=1)

# 2. small case
data, group = small_case(11, 400, 13, 19)
levels, gid = oracle.group_levels(group)
ref = pkg.synth.random_mask(400, 90, seed=4321)
thr = co.thresholds_for(gid, 2, 0.01)
out = co.identify_degs(data, gid, 2, thr, 1.0, 0.05, ref, 128, 5, seed=7)
np.savez_compressed(os.path.join(HERE, "small_case.npz"), data=data, gid=gid, ref=ref, thr=thr,
                    result=out["result"], updown=out["updown"], final_ref=out["final_ref"],
                    iters=np.array(out["iters"]), seed=7)

# 3. bundled data
ref_dir = "/root/reference/test"
if os.path.isdir(ref_dir):
    import pandas as pd
    expr = pd.read_csv(os.path.join(ref_dir, "fn_expr.txt"), sep="\t")
    meta = pd.read_csv(os.path.join(ref_dir, "fn_meta.txt"), sep="\t")
    mat = expr.iloc[:, 1:].to_numpy().astype(np.int64)
    keep = (mat > 0).sum(axis=1) > 0  # src:626, min_features = 0
    mat = mat[keep]
    levels, gid = oracle.group_levels(list(meta.iloc[:, 1]))
    ref = pkg.synth.random_mask(mat.shape[0], 3000, seed=4321)
    thr = co.thresholds_for(gid, 2, 0.01)
    out = co.identify_degs(mat, gid, 2, thr, 1.0, 0.05, ref, 128, 5, seed=7)
    print("bundled:", mat.shape, "thr", thr.tolist(), "iters", out["iters"], "deg log", out["deg_log"][0][:8],
          "up/down", int((out["updown"] == 1).sum()), int((out["updown"] == -1).sum()))
    np.savez_compressed(os.path.join(HERE, "bundled_c1.npz"), data=mat.astype(np.int32), gid=gid, ref=ref, thr=thr,
                        tables=out["result"][0][:, 2:11].astype(np.int32), pval=out["result"][0][:, 0],
                        padj=out["result"][0][:, 1], stat=out["result"][0][:, 11:15], updown=out["updown"][0],
                        final_ref=out["final_ref"][0], iters=np.array(out["iters"]), seed=7)
