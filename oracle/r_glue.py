"""CPU restatement of the R code either side of the native loop.  TEST INFRASTRUCTURE ONLY.

Follows, function by function (paths relative to /root/reference):
  prepare()                      R/core.R:269-322 (reorder), :340-341 (degrees), :345-374
                                 (threshold parsing), :383-402 (COO edge list), :407-415 (init)
  euclidean_embedding()          R/core.R:184-528 (the whole single fit, loop = cpu_oracle)
  error_calculator_comparison()  R/error_metrics.R:55-144
  make_folds()/likelihood_function()  R/adaptive_sampling.R:2552-2726

R matrices are represented as numpy arrays: float arrays with NaN for NA, or object / str
arrays whose entries may be numbers, None/NaN (NA) or strings such as "<5", ">2", "3.5", "NA".
Two declared substitutions: R's RNG (stats::runif for the initial positions, sample() for the
folds) is replaced by numpy Generator draws with the same distribution, so parity tests pass
`initial_positions` and fold index lists explicitly.
"""
from __future__ import annotations

import math

import numpy as np

from . import cpu_oracle


def _is_na(x) -> bool:
    if x is None:
        return True
    if isinstance(x, str):
        return x == "NA"
    try:
        return math.isnan(float(x))
    except (TypeError, ValueError):
        return False


def _as_numeric(x) -> float:
    """R's as.numeric on one element: unparsable strings (incl. '<5') become NA."""
    if _is_na(x):
        return math.nan
    if isinstance(x, str):
        try:
            return float(x)
        except ValueError:
            return math.nan
    return float(x)


def parse_matrix(m):
    """-> (value, code, is_na) as R/core.R:345-374: value is Inf where NA, code in {-1,0,1}."""
    m = np.asarray(m)
    n = m.shape[0]
    value = np.full((n, n), np.inf)
    code = np.zeros((n, n), dtype=np.int32)
    is_na = np.zeros((n, n), dtype=bool)
    for i in range(n):
        for j in range(n):
            x = m[i, j]
            if _is_na(x):
                is_na[i, j] = True
                continue
            if isinstance(x, str) and x.startswith(">"):
                code[i, j] = 1
                value[i, j] = _as_numeric(x[1:])
            elif isinstance(x, str) and x.startswith("<"):
                code[i, j] = -1
                value[i, j] = _as_numeric(x[1:])
            else:
                value[i, j] = _as_numeric(x)
    return value, code, is_na


def spectral_order(m):
    """R/core.R:269-322: order(avg of row/col means of the prefix-stripped numeric matrix),
    or None when the reorder is skipped."""
    m = np.asarray(m)
    n = m.shape[0]
    num = np.full((n, n), np.nan)
    for i in range(n):
        for j in range(n):
            x = m[i, j]
            if _is_na(x) or i == j:
                continue
            if isinstance(x, str) and x[:1] in "<>":
                x = x[1:]
            num[i, j] = _as_numeric(x)
    with np.errstate(invalid="ignore"):
        cnt_r = np.sum(~np.isnan(num), axis=1)
        cnt_c = np.sum(~np.isnan(num), axis=0)
        row_means = np.where(cnt_r > 0, np.nansum(num, axis=1) / np.maximum(cnt_r, 1), np.nan)
        col_means = np.where(cnt_c > 0, np.nansum(num, axis=0) / np.maximum(cnt_c, 1), np.nan)
    avg = (row_means + col_means) / 2
    avg[np.isnan(avg)] = 0
    if np.sum(avg > 0) > 1:
        return np.argsort(avg, kind="stable")
    return None


def prepare(dissimilarity_matrix, preserve_order=False):
    """Everything euclidean_embedding does before the native call, except initial positions."""
    m = np.asarray(dissimilarity_matrix)
    n = m.shape[0]
    order = None
    if n > 1 and not preserve_order:
        order = spectral_order(m)
        if order is not None:
            m = m[np.ix_(order, order)]
    value, code, is_na = parse_matrix(m)
    degrees = np.sum(~is_na, axis=1).astype(np.int32)  # rowSums(!is.na): the diagonal counts
    ei, ej = [], []
    for j in range(n):  # which(arr.ind=TRUE) walks column-major: by column j, then row i
        for i in range(j):
            v = value[i, j]
            if not math.isnan(v) and v != math.inf:
                ei.append(i)
                ej.append(j)
    ei = np.asarray(ei, dtype=np.int32)
    ej = np.asarray(ej, dtype=np.int32)
    return dict(matrix=m, order=order, n=n, degrees=degrees, edge_i=ei, edge_j=ej,
                edge_dist=value[ei, ej] if len(ei) else np.zeros(0),
                edge_thresh=code[ei, ej] if len(ei) else np.zeros(0, np.int32),
                value=value, code=code, is_na=is_na)


def initial_positions(matrix, ndim, rng):
    """R/core.R:407-415 with numpy draws: row 0 = 0, others cumulative U(0, 2*max/n) steps.
    Threshold strings are NA for the max (as.numeric(as.character()))."""
    m = np.asarray(matrix)
    n = m.shape[0]
    vals = [_as_numeric(x) for x in m.ravel()]
    vals = [v for v in vals if not math.isnan(v)]
    init_step = max(vals) / n
    steps = rng.uniform(0, 2 * init_step, size=(n - 1, ndim))
    return np.vstack([np.zeros((1, ndim)), np.cumsum(steps, axis=0)])


def dist_matrix(positions):
    """as.matrix(stats::dist(positions)): R's C loop, sum over columns then sqrt."""
    p = np.asarray(positions, dtype=np.float64)
    n, d = p.shape
    out = np.zeros((n, n))
    for c in range(d):
        diff = p[:, c][:, None] - p[:, c][None, :]
        out += diff * diff
    return np.sqrt(out)


def euclidean_embedding(dissimilarity_matrix, ndim, mapping_max_iter=1000, k0=None, cooling_rate=None,
                        c_repulsion=None, relative_epsilon=1e-4, convergence_counter=5,
                        initial_positions_=None, convergence_check_freq=3, preserve_order=False, *,
                        seed=0, pair_order=None, rng=None, dense=True):
    """Single fit, restated.  Returns the dict image of the `topolow` object (R/core.R:505-525)
    plus `order` (the row permutation applied, or None)."""
    prep = prepare(dissimilarity_matrix, preserve_order)
    if len(prep["edge_i"]) == 0:
        raise ValueError("No valid off-diagonal measurements found in dissimilarity matrix")
    m = prep["matrix"]
    if initial_positions_ is None:
        init = initial_positions(m, ndim, rng or np.random.default_rng(seed))
    else:
        init = np.asarray(initial_positions_, dtype=np.float64)
    res = cpu_oracle.optimize_layout_exact(
        init, prep["degrees"], prep["edge_i"], prep["edge_j"], prep["edge_dist"], prep["edge_thresh"],
        mapping_max_iter, k0, cooling_rate, c_repulsion, relative_epsilon, convergence_counter,
        convergence_check_freq, seed=seed, pair_order=pair_order, dense=dense)
    positions = res["positions"]
    p_dist = dist_matrix(positions)
    raw = np.array([[_as_numeric(x) for x in row] for row in m])  # thresholds -> NA (R/core.R:479)
    valid = ~np.isnan(raw)
    mae = float(np.mean(np.abs(raw[valid] - p_dist[valid]))) if valid.any() else math.nan
    return dict(positions=positions, est_distances=p_dist, mae=mae, iter=res["iterations"],
                parameters=dict(ndim=ndim, k0=k0, cooling_rate=cooling_rate, c_repulsion=c_repulsion,
                                method="cpp_exact_full_pairwise"),
                convergence=dict(achieved=res["converged"], error=res["final_mae"], final_k=res["final_k"]),
                order=prep["order"])


def error_calculator_comparison(predicted, true, input_=None):
    """R/error_metrics.R:89-143 on flattened (column-major) matrices.  Returns the report
    columns as arrays plus Completeness."""
    pred = np.asarray(predicted, dtype=np.float64)
    true_m = np.asarray(true)
    inp_m = true_m if input_ is None else np.asarray(input_)
    if pred.shape != true_m.shape or pred.shape != inp_m.shape:
        raise ValueError("All matrices must have the same dimensions")
    input_vec = np.array([_as_numeric(x) for x in inp_m.ravel(order="F")])
    truth_vec = np.array([_as_numeric(x) for x in true_m.ravel(order="F")])
    pred_vec = pred.ravel(order="F")
    missing = np.isnan(input_vec)
    pred_obs = np.where(missing, np.nan, pred_vec)
    pred_mis = np.where(missing, pred_vec, np.nan)
    in_err = truth_vec - pred_obs
    out_err = truth_vec - pred_mis
    nz = ~np.isnan(truth_vec) & (np.nan_to_num(truth_vec, nan=0.0) > 0)
    in_pct = np.full_like(in_err, np.nan)
    out_pct = np.full_like(out_err, np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        in_pct[nz] = in_err[nz] / truth_vec[nz] * 100
        out_pct[nz] = out_err[nz] / truth_vec[nz] * 100
    validation_count = int(np.sum(~np.isnan(truth_vec[missing])))
    pred_for_validation = int(np.sum(~np.isnan(out_err)))
    if validation_count > 0:
        completeness = pred_for_validation / validation_count
    else:
        total_possible = int(np.sum(~np.isnan(truth_vec)))
        total_predictions = int(np.sum(~np.isnan(pred_vec)))
        completeness = total_predictions / total_possible if total_possible > 0 else 0
    return dict(InSampleError=in_err, OutSampleError=out_err, InSamplePercentageError=in_pct,
                OutSamplePercentageError=out_pct, Completeness=completeness)


def make_folds(dissimilarity_matrix, folds, rng):
    """R/adaptive_sampling.R:2568-2598: per fold, sample holdout_size column-major linear
    indices of the still non-NA cells; NA both (r,c) and (c,r) in the pool."""
    m = np.asarray(dissimilarity_matrix)
    n = m.shape[0]
    non_na = np.array([[not _is_na(x) for x in row] for row in m])
    num_elements = int(non_na.sum())
    holdout_size = num_elements // (folds * 2)
    pool = non_na.copy()
    out = []
    for _ in range(folds):
        if int(pool.sum()) < holdout_size:
            break  # "Could not create all folds due to data sparsity. Using fewer folds."
        lin = np.flatnonzero(pool.ravel(order="F"))
        pick = rng.choice(lin, size=holdout_size, replace=False)
        out.append(pick)
        for idx in pick:
            r, c = idx % n, idx // n
            pool[r, c] = False
            pool[c, r] = False
    return out


def mask_fold(dissimilarity_matrix, holdout_indices):
    """R/adaptive_sampling.R:2608-2616."""
    m = np.array(dissimilarity_matrix, dtype=object, copy=True)
    n = m.shape[0]
    for idx in holdout_indices:
        r, c = int(idx) % n, int(idx) // n
        m[r, c] = None
        m[c, r] = None
    return m


def likelihood_function(dissimilarity_matrix, mapping_max_iter, relative_epsilon, N, k0, cooling_rate,
                        c_repulsion, folds=20, preserve_order=True, *, fold_indices=None, rng=None,
                        init_list=None, seed=0, pair_orders=None):
    """R/adaptive_sampling.R:2552-2726, sequential branch.  `fold_indices` / `init_list` inject
    the two R-RNG dependent inputs (fold hold-out index lists; per-fold initial positions); `pair_orders[f]`
    (optional, [n_iter][pairs][2]) replaces the shuffle of fold f's fit by an explicit pair order."""
    rng = rng or np.random.default_rng(seed)
    if fold_indices is None:
        fold_indices = make_folds(dissimilarity_matrix, folds, rng)
    rows = []
    for f, hold in enumerate(fold_indices):
        train = mask_fold(dissimilarity_matrix, hold)
        try:
            res = euclidean_embedding(train, N, mapping_max_iter, k0, cooling_rate, c_repulsion,
                                      relative_epsilon, 5, None if init_list is None else init_list[f],
                                      preserve_order=preserve_order, seed=seed + f, rng=rng,
                                      pair_order=None if pair_orders is None else pair_orders[f])
        except Exception:
            rows.append(dict(Holdout_MAE=math.nan, n_samples=0, sum_abs_errors=0.0, iter=math.nan, converged=0))
            continue
        err = error_calculator_comparison(res["est_distances"], dissimilarity_matrix, train)
        ose = err["OutSampleError"]
        ose = ose[~np.isnan(ose)]
        n_s = len(ose)
        s = float(np.sum(np.abs(ose)))
        rows.append(dict(Holdout_MAE=s / n_s if n_s > 0 else math.nan, n_samples=n_s, sum_abs_errors=s,
                         iter=res["iter"], converged=int(res["convergence"]["achieved"])))
    valid = [r for r in rows if not math.isnan(r["Holdout_MAE"])]
    if not valid:
        return dict(Holdout_MAE=math.nan, NLL=math.nan, mean_iter=math.nan, pct_converged=math.nan, folds=rows)
    total_samples = sum(r["n_samples"] for r in valid)
    total_abs = sum(r["sum_abs_errors"] for r in valid)
    pooled = total_abs / total_samples if total_samples > 0 else math.nan
    nll = total_samples * (1 + math.log(2 * pooled)) if not math.isnan(pooled) else math.nan
    return dict(Holdout_MAE=pooled, NLL=nll, mean_iter=float(np.mean([r["iter"] for r in valid])),
                pct_converged=float(np.mean([r["converged"] for r in valid]) * 100), folds=rows)


# ---------------------------------------------------------------------------------------------
# R/utils.R: create_cv_folds (:69-150), check_matrix_connectivity (:199-253),
# subsample_dissimilarity_matrix (:321-463); R/diagnostics.R:434-472 (adjacency, completeness)
# ---------------------------------------------------------------------------------------------
def create_cv_folds(dissimilarity_matrix, n_folds, picks_per_fold):
    """R/utils.R:103-147 with the sample() draws injected: `picks_per_fold[f]` are the 1-based positions, in
    which(!is.na(sampling_pool)) order (column-major), is replaced by explicit 0-based linear indices."""
    m = np.asarray(dissimilarity_matrix)
    nrow, ncol = m.shape
    pool = np.array([[not _is_na(x) for x in row] for row in m])
    out = []
    for f in range(n_folds):
        train = np.array(m, dtype=object, copy=True)
        for index in picks_per_fold[f]:                      # `index` here is 0-based: (index) %/% nrow, (index) %% ncol
            row, col = int(index) // nrow, int(index) % ncol
            train[row, col] = None
            train[col, row] = None
            pool[row, col] = False
            pool[col, row] = False
        out.append(train)
    return out, pool


def check_matrix_connectivity(dissimilarity_matrix):
    """R/utils.R:199-253 + R/diagnostics.R:444-464.  igraph::components is replaced by scipy's connected_components
    (declared substitution: only the component count is used)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    m = np.asarray(dissimilarity_matrix)
    n = m.shape[0]
    adj = np.array([[not _is_na(x) for x in row] for row in m])
    for i in range(n):
        adj[i, i] = False
    ncomp, _ = connected_components(csr_matrix(adj | adj.T), directed=False)
    return dict(is_connected=ncomp == 1, n_components=int(ncomp), completeness=adj.sum() / (n * (n - 1)), n_points=n,
                n_measurements=adj.sum() / 2)


def subsample_dissimilarity_matrix(dissimilarity_matrix, attempts):
    """R/utils.R:371-440: the attempts one after another (`attempts[k]` = the indices sample() returned in attempt k,
    already sorted when preserve_order); returns (attempt number, indices, connectivity) of the first connected one,
    or None."""
    m = np.asarray(dissimilarity_matrix)
    for k, sel in enumerate(attempts):
        sub = m[np.ix_(sel, sel)]
        c = check_matrix_connectivity(sub)
        if c["is_connected"]:
            return k + 1, sel, c
    return None


# ---------------------------------------------------------------------------------------------
# The sampler: R/adaptive_sampling.R weighted_kde (:1901-1935), calculate_weighted_marginals (:2457-2519),
# generate_kde_samples (:1804-1885); R/data_preprocessing.R clean_data / detect_outliers_mad (:864-879, :956-996)
# ---------------------------------------------------------------------------------------------
def _median(v):
    s = sorted(v)
    k = len(s)
    return s[k // 2] if k % 2 else 0.5 * (s[k // 2 - 1] + s[k // 2])


def clean_data(x, k=3):
    vals = [v for v in x if not math.isnan(v)]
    med = _median(vals)
    mad = 1.4826 * _median([abs(v - med) for v in vals])
    return [math.nan if (not math.isnan(v) and abs(v - med) > k * mad) else v for v in x]


def weighted_kde(x, weights, n=512):
    tot = sum(weights)
    w = [a / tot for a in weights]
    m = sum(x) / len(x)
    sd = math.sqrt(sum((a - m) ** 2 for a in x) / (len(x) - 1))
    bw = 1.06 * sd * len(x) ** (-1 / 5)
    lo, hi = min(x) - bw, max(x) + bw
    by = (hi - lo) / (n - 1)
    pts = [lo + i * by for i in range(n)]
    dens = []
    for z in pts:                       # compute_density(z) = sum(weights * dnorm(z, mean = x, sd = bw))
        dens.append(sum(wi * math.exp(-0.5 * ((z - xi) / bw) ** 2) / (bw * math.sqrt(2 * math.pi)) for wi, xi in zip(w, x)))
    return dict(x=pts, y=dens)


def _weights(score, temperature=0.1):
    lo, hi = min(score), max(score)
    w = [math.exp(-((s - lo) / (hi - lo + 1e-10)) / temperature) for s in score]
    tot = sum(w)
    return [a / tot for a in w]


def _filter_mae(table):
    keep = [i for i, v in enumerate(table["Holdout_MAE"]) if not math.isnan(v) and v > 0]
    t = {k: [table[k][i] for i in keep] for k in table}
    t["Holdout_MAE"] = clean_data(t["Holdout_MAE"], 3)
    keep = [i for i, v in enumerate(t["Holdout_MAE"]) if not math.isnan(v) and v > 0]
    return {k: [t[k][i] for i in keep] for k in t}


def calculate_weighted_marginals(table):
    t = _filter_mae({k: [float(v) for v in vals] for k, vals in table.items()})
    w = _weights([math.log(v) for v in t["Holdout_MAE"]])
    return {v: weighted_kde(t[v], w) for v in ("log_N", "log_k0", "log_cooling_rate", "log_c_repulsion")}


def approx_rule2(x, y, xout):
    """stats::approx(x, y, xout, rule = 2) with the default ties = mean."""
    groups = {}
    for a, b in zip(x, y):
        groups.setdefault(a, []).append(b)
    ux = sorted(groups)
    uy = [sum(groups[a]) / len(groups[a]) for a in ux]
    out = []
    for v in xout:
        if v <= ux[0]:
            out.append(uy[0])
        elif v >= ux[-1]:
            out.append(uy[-1])
        else:
            k = max(i for i in range(len(ux)) if ux[i] <= v)
            out.append(uy[k] + (uy[k + 1] - uy[k]) * (v - ux[k]) / (ux[k + 1] - ux[k]))
    return out


def generate_kde_samples(table, n, uniforms):
    """`uniforms[param]` = the n runif draws of the inverse-transform step (the epsilon draw changes nothing)."""
    t = {k: [float(v) for v in vals] for k, vals in table.items()}
    t["Holdout_MAE"] = clean_data(t["Holdout_MAE"], 3)
    keep = [i for i, v in enumerate(t["Holdout_MAE"]) if not math.isnan(v) and v > 0]
    t = {k: [t[k][i] for i in keep] for k in t}
    w = _weights([-math.log(v) for v in t["Holdout_MAE"]])
    out = {}
    for param in ("log_N", "log_k0", "log_cooling_rate", "log_c_repulsion"):
        kde = weighted_kde(t[param], w)
        tot = sum(kde["y"])
        cdf, run = [], 0.0
        for v in kde["y"]:
            run += v
            cdf.append(run / tot)
        out[param] = approx_rule2(cdf, kde["x"], uniforms[param])
    return out
