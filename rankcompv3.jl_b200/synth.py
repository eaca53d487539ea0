"""Synthetic expression matrices of the shapes BASELINE.json names.  numpy default_rng;
data seed 1234, reference-mask seed 4321, tie seed 7."""
from __future__ import annotations

import numpy as np

DATA_SEED = 1234
MASK_SEED = 4321
TIE_SEED = 7


def _de_structure(rng, r, frac_de=0.10):
    n_de = int(round(r * frac_de))
    de_idx = rng.choice(r, n_de, replace=False)
    fold = rng.uniform(2.0, 4.0, n_de)
    up = np.zeros(n_de, dtype=bool)
    up[: n_de // 2] = True
    rng.shuffle(up)
    lfc = np.zeros(r)
    lfc[de_idx] = np.where(up, np.log(fold), -np.log(fold))
    return lfc


def _nb(rng, mu, disp):
    lam = rng.gamma(shape=1.0 / disp, scale=mu * disp)
    return rng.poisson(lam).astype(np.int64)


def bulk(r=20000, n1=100, n2=100, seed=DATA_SEED):
    """Config 2: negative-binomial bulk RNA-seq counts, gene log-mean ~ N(4.6, 1.8^2), dispersion 0.2, 10 % DE."""
    rng = np.random.default_rng(seed)
    logmu = rng.normal(4.6, 1.8, r)
    lfc = _de_structure(rng, r)
    mu1 = np.exp(logmu)[:, None] * np.ones((1, n1))
    mu2 = np.exp(logmu + lfc)[:, None] * np.ones((1, n2))
    data = np.concatenate([_nb(rng, mu1, 0.2), _nb(rng, mu2, 0.2)], axis=1)
    group = ["group1"] * n1 + ["group2"] * n2
    return data, group, lfc != 0


def scrna(r=30000, n1=10000, n2=10000, seed=DATA_SEED, chunk=2000):
    """Configs 3-5: zero-inflated NB single-cell counts, gene log-mean ~ N(-1.5, 1.5^2), ~85-90 % zeros."""
    rng = np.random.default_rng(seed)
    logmu = rng.normal(-1.5, 1.5, r)
    lfc = _de_structure(rng, r)
    pz = rng.uniform(0.0, 0.3, r)  # extra dropout per gene
    data = np.empty((r, n1 + n2), dtype=np.int64)
    for (c0, n, lm) in ((0, n1, logmu), (n1, n2, logmu + lfc)):
        for s0 in range(0, n, chunk):
            m = min(chunk, n - s0)
            mu = np.exp(lm)[:, None] * np.ones((1, m))
            x = _nb(rng, mu, 0.5)
            x[rng.random((r, m)) < pz[:, None]] = 0
            data[:, c0 + s0:c0 + s0 + m] = x
    group = ["group1"] * n1 + ["group2"] * n2
    return data, group, lfc != 0


def tie_free(data, seed=DATA_SEED):
    """Tie-free twin: per-sample dense order of the data with ties broken by a seeded jitter, so every
    column is a permutation of distinct integers (the real reference is deterministic on such input)."""
    rng = np.random.default_rng(seed + 1)
    r, c = data.shape
    out = np.empty((r, c), dtype=np.int64)
    for s in range(c):
        key = data[:, s].astype(np.float64) + rng.random(r) * 0.5
        out[np.argsort(key, kind="stable"), s] = np.arange(r)
    return out


def reference_mask(is_de, n_ref=3000, seed=MASK_SEED):
    """Seeded 'house-keeping' reference set: n_ref non-DE genes."""
    rng = np.random.default_rng(seed)
    cand = np.nonzero(~np.asarray(is_de))[0]
    pick = rng.choice(cand, min(n_ref, len(cand)), replace=False)
    m = np.zeros(len(is_de), dtype=bool)
    m[pick] = True
    return m


def random_mask(r, n_ref=3000, seed=MASK_SEED):
    """src:635: a random sample of n_ref genes (what reoa falls back to on the bundled data)."""
    rng = np.random.default_rng(seed)
    m = np.zeros(r, dtype=bool)
    m[rng.choice(r, min(n_ref, r), replace=False)] = True
    return m


def scrna_torch(r=30000, n1=10000, n2=10000, seed=DATA_SEED, device="cuda:0", chunk=1000):
    """Same model as scrna() but generated on the GPU with torch (seconds instead of minutes at 30k x 20k).
    Returns (int64 tensor [c, r] row-major == column-major r x c, group list, is_de mask)."""
    import torch
    rng = np.random.default_rng(seed)
    logmu = rng.normal(-1.5, 1.5, r)
    lfc = _de_structure(rng, r)
    pz = rng.uniform(0.0, 0.3, r)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    torch.manual_seed(seed)   # _standard_gamma draws from the default generator: make the matrix reproducible
    out = torch.empty((n1 + n2, r), dtype=torch.int64, device=device)
    pz_t = torch.as_tensor(pz, device=device, dtype=torch.float32)
    for (c0, n, lm) in ((0, n1, logmu), (n1, n2, logmu + lfc)):
        mu = torch.as_tensor(np.exp(lm), device=device, dtype=torch.float32)
        for s0 in range(0, n, chunk):
            m = min(chunk, n - s0)
            # NB(mu, dispersion 0.5) as a gamma-Poisson mixture: gamma(shape 2, scale mu/2)
            gam = torch._standard_gamma(torch.full((m, r), 2.0, device=device)) * (mu * 0.5)
            x = torch.poisson(gam, generator=g)
            x = torch.where(torch.rand((m, r), device=device, generator=g) < pz_t, torch.zeros_like(x), x)
            out[c0 + s0:c0 + s0 + m] = x.to(torch.int64)
    group = ["group1"] * n1 + ["group2"] * n2
    return out, group, lfc != 0
