"""Synthetic low-rank dissimilarity problems, in edge-list form.

X in R^{n x r} (r = ndim): 4 Gaussian clusters along a linear trend; D_ij = ||x_i - x_j|| * (1 + 0.05 eps),
eps ~ N(0,1), clipped at 0.1; each upper-triangle pair kept independently with probability (1 - missing);
a random spanning path keeps the measured graph connected; 5 % of kept entries become '>' thresholds at
the 90th percentile and 5 % '<' at the 10th.  The diagonal is measured (0), so degrees include it
(R/core.R:340-341).  Never materialises n x n: usable at n = 100k.
"""
from __future__ import annotations

import numpy as np


def _unrank_upper(lin, n):
    """Linear index over the upper triangle (row-major, i<j) -> (i, j), vectorised, exact."""
    lin = lin.astype(np.int64)
    nn = n - 0.5
    i = np.floor(nn - np.sqrt(np.maximum(nn * nn - 2.0 * lin, 0.0))).astype(np.int64)
    i = np.clip(i, 0, n - 2)
    start = i * (2 * n - i - 1) // 2
    i = np.where(start > lin, i - 1, i)
    start = i * (2 * n - i - 1) // 2
    nxt = (i + 1) * (2 * n - i - 2) // 2
    i = np.where(nxt <= lin, i + 1, i)
    start = i * (2 * n - i - 1) // 2
    j = lin - start + i + 1
    return i, j


def make_problem(n, ndim, missing, seed=0, thresholds=True, init_scale=None):
    """-> dict(initial_positions, degrees, edge_i, edge_j, edge_dist, edge_thresh, X)."""
    rng = np.random.default_rng(seed)
    centers = rng.normal(size=(4, ndim)) * 4.0
    trend = rng.normal(size=ndim)
    trend /= np.linalg.norm(trend)
    which = rng.integers(0, 4, size=n)
    t = np.sort(rng.uniform(0, 10, size=n))
    X = (centers[which] + rng.normal(size=(n, ndim)) + t[:, None] * trend[None, :]).astype(np.float64)

    P = n * (n - 1) // 2
    p = 1.0 - missing
    # geometric skipping over the linear pair index
    est = int(P * p * 1.02 + 1000)
    lin = np.cumsum(rng.geometric(p, size=est).astype(np.int64)) - 1
    while lin[-1] < P - 1 and p < 1.0:
        more = np.cumsum(rng.geometric(p, size=max(est // 50, 1000)).astype(np.int64)) + lin[-1]
        lin = np.concatenate([lin, more])
    lin = lin[lin < P]
    ei, ej = _unrank_upper(lin, n)
    # spanning path over a random permutation of the points
    perm = rng.permutation(n)
    a, b = np.minimum(perm[:-1], perm[1:]), np.maximum(perm[:-1], perm[1:])
    key = np.concatenate([ei * n + ej, a.astype(np.int64) * n + b])
    key = np.unique(key) if len(key) < 5_000_000 else _unique_sorted_merge(ei * n + ej, a.astype(np.int64) * n + b)
    ei, ej = (key // n).astype(np.int32), (key % n).astype(np.int32)
    E = len(ei)

    ed = np.empty(E, dtype=np.float64)
    Xf = X.astype(np.float32)
    step = 4_000_000
    for s in range(0, E, step):
        df = Xf[ei[s:s + step]] - Xf[ej[s:s + step]]
        ed[s:s + step] = np.sqrt(np.einsum("ij,ij->i", df, df))
    noise = rng.standard_normal(E).astype(np.float32)
    ed = np.maximum(ed * (1.0 + 0.05 * noise), 0.1)
    et = np.zeros(E, dtype=np.int32)
    if thresholds:
        sample = ed[:: max(1, E // 200_000)]
        hi, lo = np.quantile(sample, 0.9), np.quantile(sample, 0.1)
        u = rng.random(E)
        gt = u < 0.05
        lt = (u >= 0.05) & (u < 0.10)
        et[gt], et[lt] = 1, -1
        ed[gt], ed[lt] = hi, lo
    deg = (np.bincount(ei, minlength=n) + np.bincount(ej, minlength=n) + 1).astype(np.int32)
    # R/core.R:407-415 initialisation: cumulative U(0, 2*max/n) steps from the origin
    init_step = (ed[et == 0].max() if init_scale is None else init_scale) / n
    init = np.vstack([np.zeros((1, ndim)), np.cumsum(rng.uniform(0, 2 * init_step, size=(n - 1, ndim)), axis=0)])
    return dict(initial_positions=init, degrees=deg, edge_i=ei, edge_j=ej, edge_dist=ed, edge_thresh=et, X=X)


def _unique_sorted_merge(sorted_keys, extra):
    """Union of an already sorted unique key array with a small extra set."""
    extra = np.unique(extra)
    pos = np.searchsorted(sorted_keys, extra)
    pos_c = np.minimum(pos, len(sorted_keys) - 1)
    new = extra[sorted_keys[pos_c] != extra]
    if len(new) == 0:
        return sorted_keys
    out = np.concatenate([sorted_keys, new])
    out.sort(kind="stable")
    return out


def fit_args(prob):
    """Positional arguments (initial_positions ... edge_thresh) of _lib.fit / _lib.Plan / the oracle."""
    return (prob["initial_positions"], prob["degrees"], prob["edge_i"], prob["edge_j"], prob["edge_dist"],
            prob["edge_thresh"])


def dense_matrix(prob, n):
    """The R-style matrix (object array with '<x' / '>x' strings, NaN = NA, 0 diagonal) for small n."""
    m = np.full((n, n), np.nan, dtype=object)
    for i in range(n):
        m[i, i] = 0.0
    for a, b, v, t in zip(prob["edge_i"], prob["edge_j"], prob["edge_dist"], prob["edge_thresh"]):
        s = float(v) if t == 0 else ((">" if t > 0 else "<") + repr(float(v)))
        m[a, b] = s
        m[b, a] = s
    return m
