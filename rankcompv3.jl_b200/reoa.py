"""
reoa(): host driver with the reference's keyword API (src/RankCompV3.jl:536-685).

In a deployment the Julia host keeps this function unchanged and only src:652-662 is re-pointed at
libreo_cuda.so.  Julia is not available in this image, so this module mirrors the
driver in Python -- file I/O, meta checks (src:563-606), pseudo-bulk (src:56-67, 608-612), filters
(src:618-628), reference-gene selection (src:635-651), TSV outputs (src:663-683) -- and calls the same
C ABI.  Plotting (code/plot.jl) is out of scope.  Reference quirks are kept: yes/no flags are strings,
`expr_threshold` is accepted and unused, cells are filtered by `> min_profiles` and genes by
`> min_features` (src:618, 626), columns are matched to meta rows by position (src:614-615).
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from . import api

_HK_DEFAULT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hk_gene_file", "HK_genes_info.tsv")


def pseudobulk_group(group_expr: np.ndarray, n_pseudo: int, rng: np.random.Generator):
    """src:56-67: shuffle the cells of one group, partition into chunks of ceil(c/n_pseudo), sum rows."""
    r, c = group_expr.shape
    cp = int(np.ceil(c / n_pseudo))
    perm = rng.permutation(c)
    parts = [perm[i:i + cp] for i in range(0, c, cp)]
    return np.stack([group_expr[:, p].sum(axis=1) for p in parts], axis=1)


def reoa(fn_expr: str = "fn_expr.txt", fn_metadata: str = "fn_metadata.txt", *, expr_threshold=0,
         min_profiles: int = 0, min_features: int = 0, pval_reo: float = 0.01, pval_deg: float = 1.0,
         padj_deg: float = 0.05, n_pseudo: int = 0, use_hk_genes: str = "yes", hk_file: str = _HK_DEFAULT,
         gene_name_type: str = "ENSEMBL", ref_gene_max: int = 3000, ref_gene_min: int = 100, n_iter: int = 128,
         n_conv: int = 5, work_dir: str = "./", use_testdata: str = "no", seed: int = 0, testdata_dir: str | None = None,
         handle: api.Reo | None = None, write_files: bool = True):
    """Returns the gene_up_down DataFrame (src:682-684).  `seed` (new, defaulted) keys the host-side random
    choices (pseudo-bulk shuffle, random reference sample) and the tie coins."""
    work_dir = os.path.abspath(work_dir)
    if use_testdata == "yes":  # src:558-561
        if testdata_dir is None:
            raise ValueError("ArgumentError: use_testdata='yes' needs testdata_dir (the reference's test/ directory)")
        fn_expr = os.path.join(testdata_dir, "fn_expr.txt")
        fn_metadata = os.path.join(testdata_dir, "fn_meta.txt")
    else:
        fn_expr = fn_expr if os.path.isabs(fn_expr) else os.path.join(work_dir, fn_expr)
        fn_metadata = fn_metadata if os.path.isabs(fn_metadata) else os.path.join(work_dir, fn_metadata)
    if not (os.path.isfile(fn_expr) and os.path.isfile(fn_metadata)):
        raise ValueError(f"ArgumentError: {fn_expr}, or {fn_metadata}, does not exist or is not a regular file.")
    if not (os.path.getsize(fn_expr) > 0 and os.path.getsize(fn_metadata) > 0):
        raise ValueError(f"ArgumentError: {fn_expr}, or {fn_metadata}, has size 0.")
    fn_stem = os.path.splitext(os.path.basename(fn_expr))[0]
    expr = pd.read_csv(fn_expr, sep=None, engine="python")
    meta = pd.read_csv(fn_metadata, sep=None, engine="python")
    if meta.shape[1] < 2:
        raise ValueError(f"ArgumentError: {fn_metadata} the file for meta data, has only 0 or 1 column.")
    if not {"Name", "Group"} <= set(meta.columns) and set(meta.iloc[:, 0]) <= set(expr.columns):
        meta = meta.rename(columns={meta.columns[0]: "Name", meta.columns[1]: "Group"})
    if not {"Name", "Group"} <= set(meta.columns) or not set(meta["Name"]) <= set(expr.columns):
        raise ValueError(f"ArgumentError: Meta data file, {fn_metadata}, does not fit with the expression file, "
                         f"{fn_expr}. Some sample names in the meta are not found in the column names of the "
                         "expression matrix")
    if len(set(expr.columns)) != len(expr.columns):
        raise ValueError("ArgumentError: Duplicate column names exist in the representation matrix.")
    meta["Group"] = meta["Group"].astype(str)
    g_name = list(dict.fromkeys(meta["Group"]))
    mg = len(g_name)
    if mg < 2:
        raise ValueError(f"ArgumentError: Meta data file, {fn_metadata} has only 0 or 1 group. "
                         "It must consist of two 'Group' levels")
    if "Name" not in expr.columns and expr.columns[0] not in set(meta["Name"]):
        expr = expr.rename(columns={expr.columns[0]: "Name"})
    expr = expr.dropna()
    num_cols = [c for c in expr.columns if pd.api.types.is_numeric_dtype(expr[c])]
    if not set(meta["Name"]) <= set(num_cols):
        raise ValueError(f"ArgumentError: {fn_expr} expression matrix contains non-numeric (Number) profiles.")
    gene_names = expr["Name"].astype(str).to_numpy()
    rng = np.random.default_rng(seed)
    h = handle or api.default_handle(seed)
    if n_pseudo > 0:  # src:608-612; the shuffle/partition is decided here, the row sums run on the device
        all_cols = list(meta["Name"])
        full = expr[all_cols].to_numpy()
        col_index = {n: i for i, n in enumerate(all_cols)}
        profiles, names, groups = [], [], []
        for g in g_name:
            cols = np.array([col_index[n] for n in meta["Name"][meta["Group"] == g]], dtype=np.int32)
            cp = int(np.ceil(len(cols) / n_pseudo))  # src:60
            perm = cols[rng.permutation(len(cols))]
            parts = [perm[i:i + cp] for i in range(0, len(perm), cp)]
            names += [f"{g}_x{i + 1}" for i in range(len(parts))]
            groups += [g] * len(parts)
            profiles += parts
        mat, _ = h.pseudobulk(full, profiles)
        meta_group = pd.DataFrame({"Name": names, "Group": groups})
    else:  # src:614-615: every column after the first, matched to meta rows by position
        mat = expr.iloc[:, 1:].to_numpy()
        names = list(expr.columns[1:])
        meta_group = meta.copy()
    # src:618-628
    per_cell, _ = h.detect_counts(mat)            # device: detected genes per profile
    s_inds = per_cell > min_profiles
    if (~s_inds).any():
        dropped = {names[i] for i in np.nonzero(~s_inds)[0]}
        meta_group = meta_group[~meta_group.iloc[:, 0].isin(dropped)]
    mat = mat[:, s_inds]
    names = [n for n, keep in zip(names, s_inds) if keep]
    _, per_gene = h.detect_counts(mat)            # device: detecting profiles per gene
    inds = per_gene > min_features
    gene_names = gene_names[inds]
    mat = mat[inds, :]
    print(f"INFO: size after filtering lowly expressed genes and profiles and pseudo-bulk sampling, {mat.shape}")
    # src:635-651
    r = len(gene_names)
    ref_gene = set(gene_names[rng.choice(r, min(r, ref_gene_max), replace=False)])
    if use_hk_genes == "yes":
        if not os.path.isfile(hk_file):
            raise ValueError(f"ArgumentError: {hk_file} does not exist or is not a regular file.")
        if not os.path.getsize(hk_file) > 0:
            raise ValueError(f"ArgumentError: {hk_file} for house-keeping genes has size 0.")
        hk = pd.read_csv(hk_file, sep="\t", dtype=str)
        if gene_name_type in hk.columns:
            hk_set = set(hk[gene_name_type].dropna()) & set(gene_names)
            if len(hk_set) < ref_gene_min:
                print(f"WARN: only {len(hk_set)} house-keeping genes are available, we just ignore this.")
            else:
                ref_gene = hk_set
    ref_gene_vec = np.array([g in ref_gene for g in gene_names], dtype=bool)
    group = list(meta_group["Group"])
    res = api.identify_degs(mat, group, gene_names, pval_reo, pval_deg, padj_deg, ref_gene_vec, n_iter, n_conv,
                            handle=h)
    # src:663-683
    K = (res.shape[1] - 1) // 16
    cols = {}
    for i in range(K):
        block = pd.DataFrame(res[:, 1 + 16 * i:17 + 16 * i], columns=api.HEADER)
        block.insert(0, "genename", res[:, 0])
        fg_name = "_".join([g_name[0], g_name[1]]) if mg == 2 else g_name[i]
        if write_files:
            block.to_csv(os.path.join(work_dir, "_".join([fn_stem, fg_name, "result.tsv"])), sep="\t", index=False)
        cols[f"{g_name[0]}_vs_{g_name[1]}" if mg == 2 else f"{g_name[i]}_vs_other"] = res[:, 16 * (i + 1)]
    if write_files:
        df_expr = pd.DataFrame(mat, columns=names)
        df_expr.insert(0, "genename", gene_names)
        df_expr.to_csv(os.path.join(work_dir, fn_stem + "_df_expr.tsv"), sep="\t", index=False)
        meta_group.to_csv(os.path.join(work_dir, fn_stem + "_df_meta.tsv"), sep="\t", index=False)
    gene_up_down = pd.DataFrame({"gene_name": gene_names, **cols})
    if write_files:
        gene_up_down.to_csv(os.path.join(work_dir, fn_stem + "_gene_up_down.tsv"), sep="\t", index=False)
    return gene_up_down
