"""The batched cross-validation evaluator: mirror of likelihood_function()
(/root/reference/R/adaptive_sampling.R:2552-2726) and of the residual part of
error_calculator_comparison() (/root/reference/R/error_metrics.R:89-143) that it consumes.

The reference runs the `folds` fits of one parameter sample one after another (or forks them with
mclapply); here all folds of all samples handed to `likelihood_batch` go to the GPU in ONE
topolow_fit_batch call - the fork boundary that would break CUDA is gone.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

from . import _lib
from .core import build_problem, parse_dissimilarity, random_initial_positions


def error_calculator_comparison(predicted_dissimilarities, true_dissimilarities, input_dissimilarities=None):
    """R/error_metrics.R:55-144.  Returns report_df columns (flattened column-major, NaN = NA)
    and Completeness."""
    if not isinstance(predicted_dissimilarities, np.ndarray) or not isinstance(true_dissimilarities, np.ndarray) \
            or predicted_dissimilarities.ndim != 2 or true_dissimilarities.ndim != 2:
        raise ValueError("predicted_dissimilarities and true_dissimilarities must be matrices")
    if input_dissimilarities is None:
        input_dissimilarities = true_dissimilarities
    if not isinstance(input_dissimilarities, np.ndarray) or input_dissimilarities.ndim != 2:
        raise ValueError("input_dissimilarities must be a matrix")
    if predicted_dissimilarities.shape != true_dissimilarities.shape or \
            predicted_dissimilarities.shape != input_dissimilarities.shape:
        raise ValueError("All matrices must have the same dimensions")

    def as_numeric(m):  # threshold strings and NA both become NA
        v, c, na = parse_dissimilarity(m)
        return np.where(na | (c != 0), np.nan, v).ravel(order="F")

    input_vec = as_numeric(input_dissimilarities)
    truth_vec = as_numeric(true_dissimilarities)
    pred_vec = np.asarray(predicted_dissimilarities, dtype=np.float64).ravel(order="F")
    missing = np.isnan(input_vec)
    in_err = truth_vec - np.where(missing, np.nan, pred_vec)
    out_err = truth_vec - np.where(missing, pred_vec, np.nan)
    nz = ~np.isnan(truth_vec) & (np.nan_to_num(truth_vec, nan=0.0) > 0)
    in_pct = np.full_like(in_err, np.nan)
    out_pct = np.full_like(out_err, np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        in_pct[nz] = in_err[nz] / truth_vec[nz] * 100
        out_pct[nz] = out_err[nz] / truth_vec[nz] * 100
    validation_count = int(np.sum(~np.isnan(truth_vec[missing])))
    if validation_count > 0:
        completeness = int(np.sum(~np.isnan(out_err))) / validation_count
    else:
        total_possible = int(np.sum(~np.isnan(truth_vec)))
        completeness = int(np.sum(~np.isnan(pred_vec))) / total_possible if total_possible > 0 else 0
    return dict(report_df=dict(InSampleError=in_err, OutSampleError=out_err, InSamplePercentageError=in_pct,
                               OutSamplePercentageError=out_pct),
                Completeness=completeness)


def make_folds(dissimilarity_matrix, folds, rng):
    """R/adaptive_sampling.R:2568-2598.  Returns a list of column-major linear index arrays."""
    _, _, is_na = parse_dissimilarity(dissimilarity_matrix)
    n = is_na.shape[0]
    pool = ~is_na
    holdout_size = int(pool.sum()) // (folds * 2)
    out = []
    for _ in range(folds):
        if int(pool.sum()) < holdout_size:
            warnings.warn("Could not create all folds due to data sparsity. Using fewer folds.")
            break
        lin = np.flatnonzero(pool.ravel(order="F"))
        pick = rng.choice(lin, size=holdout_size, replace=False)
        out.append(pick)
        r, c = pick % n, pick // n
        pool[r, c] = False
        pool[c, r] = False
    return out


def _fold_job(value, code, is_na, holdout, preserve_order):
    """Training problem of one fold (R/adaptive_sampling.R:2608-2616) + its held-out cells."""
    n = value.shape[0]
    r, c = holdout % n, holdout // n
    na_train = is_na.copy()
    na_train[r, c] = True
    na_train[c, r] = True
    train = np.where(na_train, np.nan, value)
    # rebuild a matrix build_problem understands: thresholds survive through (value, code)
    obj = train.astype(object)
    for a, b in zip(*np.nonzero((code != 0) & ~na_train)):
        obj[a, b] = (">" if code[a, b] > 0 else "<") + repr(float(value[a, b]))
    prob = build_problem(obj if (code != 0).any() else train, preserve_order)
    # out-of-sample cells: NA (or threshold) in the training matrix, numeric in the truth
    truth_num = np.where(is_na | (code != 0), np.nan, value)
    train_num = np.where(na_train | (code != 0), np.nan, value)
    cells = np.isnan(train_num) & ~np.isnan(truth_num)
    ci, cj = np.nonzero(cells)
    return prob, ci.astype(np.int32), cj.astype(np.int32), truth_num[ci, cj]


def likelihood_batch(dissimilarity_matrix, samples, mapping_max_iter, relative_epsilon, folds=20,
                     preserve_order=True, *, fold_indices=None, init_list=None, rng=None, seed=0, device=0,
                     precision="f32"):
    """Evaluate many parameter samples at once.  `samples` is a list of dicts with keys N, k0,
    cooling_rate, c_repulsion; returns one likelihood_function() result per sample.  All
    len(samples) x folds fits run in a single topolow_fit_batch call.

    The folds are drawn ONCE for the whole batch (or taken from `fold_indices`): every sample is scored on
    the same hold-out cells, and topolow_fit_batch builds the device records of a fold once.  The
    reference draws fresh folds inside every likelihood_function call (R/adaptive_sampling.R:2568-2598);
    the distribution of a sample's score is the same, the scores of different samples are no longer
    independent (common random numbers - which is what comparing samples wants).  Call
    likelihood_function per sample for the reference's behaviour.

    preserve_order defaults to True here because fold residuals are aligned by position; the
    reference re-aligns by row names (R/error_metrics.R:76-87) which an unnamed matrix lacks."""
    rng = rng or np.random.default_rng(seed)
    value, code, is_na = parse_dissimilarity(dissimilarity_matrix)
    if fold_indices is None:
        fold_indices = make_folds(dissimilarity_matrix, folds, rng)
    prec_c = {"f32": _lib.PREC_F32, "f64": _lib.PREC_F64_EXACT}[precision]
    fold_jobs = [_fold_job(value, code, is_na, np.asarray(h), preserve_order) for h in fold_indices]
    # hold-out cells in the row numbering of the fold's training problem: they are scored on the device
    # at the end of each fit (topolow_problem.holdout_*), the positions need not come back for that
    fold_holdout = []
    for prob, ci, cj, tr in fold_jobs:
        if prob["order"] is not None:
            inv = np.empty(len(prob["order"]), dtype=np.int64)
            inv[prob["order"]] = np.arange(len(prob["order"]))
            ci, cj = inv[ci], inv[cj]
        fold_holdout.append((np.ascontiguousarray(ci, dtype=np.int32), np.ascontiguousarray(cj, dtype=np.int32),
                             np.ascontiguousarray(tr, dtype=np.float64)))
    jobs, meta = [], []
    for s_idx, s in enumerate(samples):
        for f_idx, (prob, ci, cj, tr) in enumerate(fold_jobs):
            ndim = int(s["N"])
            if len(prob["edge_i"]) == 0:
                meta.append((s_idx, f_idx, None))
                continue
            if init_list is not None:
                init = init_list[s_idx][f_idx]
            else:
                init = random_initial_positions(prob["value"], prob["code"], prob["is_na"], ndim, rng)
            jobs.append(dict(initial_positions=init, degrees=prob["degrees"], edge_i=prob["edge_i"],
                             edge_j=prob["edge_j"], edge_dist=prob["edge_dist"], edge_thresh=prob["edge_thresh"],
                             n_iter=int(mapping_max_iter), k0=s["k0"], cooling_rate=s["cooling_rate"],
                             c_repulsion=s["c_repulsion"], relative_epsilon=relative_epsilon, convergence_window=5,
                             precision=prec_c, seed=seed + 1000003 * s_idx + f_idx, holdout=fold_holdout[f_idx]))
            meta.append((s_idx, f_idx, len(jobs) - 1))
    results = _lib.fit_batch(jobs, device=device) if jobs else []

    per_sample = [[] for _ in samples]
    for s_idx, f_idx, j in meta:
        row = dict(Holdout_MAE=math.nan, n_samples=0, sum_abs_errors=0.0, iter=math.nan, converged=0)
        if j is not None and results[j].get("status", 1) == _lib.OK:
            res = results[j]
            s_abs, cnt = res["holdout_sum_abs"], res["holdout_count"]
            row = dict(Holdout_MAE=s_abs / cnt if cnt > 0 else math.nan, n_samples=cnt, sum_abs_errors=s_abs,
                       iter=res["iterations"], converged=int(res["converged"]))
        per_sample[s_idx].append(row)

    out = []
    for rows in per_sample:  # R/adaptive_sampling.R:2695-2725
        valid = [r for r in rows if not math.isnan(r["Holdout_MAE"])]
        if not valid:
            out.append(dict(Holdout_MAE=math.nan, NLL=math.nan, mean_iter=math.nan, pct_converged=math.nan, folds=rows))
            continue
        total = sum(r["n_samples"] for r in valid)
        tabs = sum(r["sum_abs_errors"] for r in valid)
        pooled = tabs / total if total > 0 else math.nan
        nll = math.nan if math.isnan(pooled) else total * (1 + (math.log(2 * pooled) if pooled > 0 else -math.inf))
        out.append(dict(Holdout_MAE=pooled, NLL=nll, mean_iter=float(np.mean([r["iter"] for r in valid])),
                        pct_converged=float(np.mean([r["converged"] for r in valid]) * 100), folds=rows))
    return out


def likelihood_function(dissimilarity_matrix, mapping_max_iter, relative_epsilon, N, k0, cooling_rate, c_repulsion,
                        folds=20, num_cores=1, preserve_order=True, **kw):
    """Single-sample form with the reference's signature (R/adaptive_sampling.R:2552-2555);
    num_cores is accepted and ignored (the folds already run concurrently on the device)."""
    return likelihood_batch(dissimilarity_matrix, [dict(N=N, k0=k0, cooling_rate=cooling_rate,
                                                        c_repulsion=c_repulsion)], mapping_max_iter,
                            relative_epsilon, folds, preserve_order, **kw)[0]
