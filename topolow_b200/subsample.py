"""Subsampling with a connectivity guarantee - the step in front of every fit of an `opt_subsample` run.

Mirrors /root/reference/R/utils.R:199-253 (`check_matrix_connectivity`), :321-463
(`subsample_dissimilarity_matrix`), :520-625 (`sanity_check_subsample`) and R/diagnostics.R:434-472
(`analyze_network_structure`): same names, arguments, result fields, messages.  The reference builds an n x n adjacency
matrix and an igraph object per attempt; here the measurement graph is an edge list that goes to the device once and
ALL attempts of a call are checked by one launch sequence (`topolow_components`, csrc/graph.cu: union-find over
(candidate, edge) pairs).  Candidates are drawn in the order the reference draws them, so the returned subsample is the
one its sequential loop would return from the same random stream.

R matrices are numpy arrays (float with NaN = NA, or object / str cells), as in core.py.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib, core


def _non_na(matrix):
    m = np.asarray(matrix)
    if m.dtype.kind in "fiub":
        return ~np.isnan(m.astype(np.float64))
    return ~core.parse_dissimilarity(m)[2]


def measurement_edges(matrix):
    """Upper-triangle pairs with a non-NA cell in either orientation (adjacency = !is.na, diagonal dropped;
    graph_from_adjacency_matrix(mode = "undirected") joins a pair when either cell is set)."""
    adj = _non_na(matrix)
    adj = adj | adj.T
    ii, jj = np.nonzero(np.triu(adj, 1))
    return ii.astype(np.int32), jj.astype(np.int32), adj


def analyze_network_structure(dissimilarity_matrix):
    """R/diagnostics.R:434-472 (summary part; node-level degree and completeness as arrays)."""
    m = np.asarray(dissimilarity_matrix)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise ValueError("Input must be a square matrix")
    n = m.shape[0]
    if n < 2:
        raise ValueError("Matrix must have at least 2 rows/columns")
    adj = _non_na(m)
    np.fill_diagonal(adj, False)
    deg = adj.sum(axis=1)
    return dict(adjacency=adj, connectivity=dict(degree=deg, completeness=deg / (n - 1)),
                summary=dict(n_points=n, n_measurements=adj.sum() / 2, completeness=adj.sum() / (n * (n - 1))))


def _validate(m):
    if not isinstance(m, np.ndarray) or m.ndim != 2:
        raise ValueError("dissimilarity_matrix must be a matrix")
    if m.shape[0] != m.shape[1]:
        raise ValueError("dissimilarity_matrix must be square")


def check_matrix_connectivity(dissimilarity_matrix, min_completeness=0.1, *, device=0, components_fn=None):
    """R/utils.R:199-253.  `components_fn(n, edge_i, edge_j, masks)` stands in for the device call in CPU tests."""
    _validate(dissimilarity_matrix)
    n = dissimilarity_matrix.shape[0]
    if n < 2:
        raise ValueError("dissimilarity_matrix must have at least 2 points")
    net = analyze_network_structure(dissimilarity_matrix)
    ei, ej, _ = measurement_edges(dissimilarity_matrix)
    comp = (components_fn or _device_components(device))(n, ei, ej, None)[0]
    n_components = int(comp[0])
    is_connected = n_components == 1
    completeness = float(net["summary"]["completeness"])
    if is_connected and completeness < min_completeness:
        warnings.warn("Network is connected but sparse (%.1f%% complete). This may lead to poor optimization. "
                      "Consider using more data points." % (completeness * 100))
    return dict(is_connected=is_connected, n_components=n_components, completeness=completeness,
                n_points=n, n_measurements=float(net["summary"]["n_measurements"]))


def _device_components(device):
    return lambda n, ei, ej, masks: _lib.components(n, ei, ej, masks, device)


def subsample_dissimilarity_matrix(dissimilarity_matrix, sample_size, max_attempts=5, min_completeness=0.1, random_seed=None,
                                   verbose=False, preserve_order=False, *, rownames=None, rng=None, device=0,
                                   components_fn=None):
    """R/utils.R:321-463.  `rng` (numpy Generator) stands in for R's sample(); `random_seed` seeds a fresh one."""
    _validate(dissimilarity_matrix)
    n = dissimilarity_matrix.shape[0]
    if not isinstance(sample_size, (int, float, np.integer, np.floating)) or isinstance(sample_size, bool) or sample_size < 2:
        raise ValueError("sample_size must be a numeric value >= 2")
    sample_size = int(np.floor(sample_size))
    fn = components_fn or _device_components(device)
    if sample_size >= n:
        c = check_matrix_connectivity(dissimilarity_matrix, min_completeness, device=device, components_fn=components_fn)
        return dict(subsampled_matrix=dissimilarity_matrix, selected_indices=np.arange(n), selected_names=rownames,
                    is_connected=c["is_connected"], n_components=c["n_components"], completeness=c["completeness"],
                    attempt_number=1)
    if random_seed is not None:
        rng = np.random.default_rng(int(random_seed))
    rng = rng or np.random.default_rng()
    ei, ej, adj = measurement_edges(dissimilarity_matrix)
    # every attempt's draw, in the reference's order; one device call checks them all
    picks = []
    for _ in range(int(max_attempts)):
        sel = rng.choice(n, size=sample_size, replace=False)
        picks.append(np.sort(sel) if preserve_order else sel)
    masks = np.zeros((len(picks), n), dtype=np.uint8)
    for a, sel in enumerate(picks):
        masks[a, sel] = 1
    comp, pts, edges = fn(n, ei, ej, masks)
    completeness = 2.0 * edges / (sample_size * (sample_size - 1))        # sum(adjacency) / (n (n - 1)) of the sub-matrix
    for a, sel in enumerate(picks):
        if int(comp[a]) == 1:
            if verbose:
                print("  [OK] Connected subsample obtained (attempt %d, size %d, %.1f%% complete)"
                      % (a + 1, sample_size, completeness[a] * 100))
            if completeness[a] < min_completeness:
                warnings.warn("Network is connected but sparse (%.1f%% complete). This may lead to poor optimization. "
                              "Consider using more data points." % (completeness[a] * 100))
            names = None if rownames is None else [rownames[i] for i in sel]
            return dict(subsampled_matrix=dissimilarity_matrix[np.ix_(sel, sel)], selected_indices=sel, selected_names=names,
                        is_connected=True, n_components=1, completeness=float(completeness[a]), attempt_number=a + 1)
        if verbose:
            print("  X Not connected (%d components, %.1f%% complete)" % (int(comp[a]), completeness[a] * 100))
    last = len(picks) - 1
    raise RuntimeError(
        "Failed to obtain a connected subsample after %d attempts.\n"
        "  Final sample size tried: %d (started at %d)\n"
        "  Original matrix size: %d\n"
        "  Last attempt had %d components with %.1f%% completeness\n\n"
        "Possible solutions:\n"
        "  1. Increase opt_subsample (current: %d)\n"
        "  2. Reduce number of CV folds\n"
        "  3. Use full dataset (opt_subsample = NULL)\n"
        "  4. Check if your data has inherent disconnected groups"
        % (max_attempts, sample_size, sample_size, n, int(comp[last]), completeness[last] * 100, sample_size))


def sanity_check_subsample(subsampled_matrix, folds=20, min_points_per_fold=3, min_measurements_per_fold=3, verbose=True):
    """R/utils.R:520-625: the five checks, the same messages, the same diagnostics."""
    n_points = subsampled_matrix.shape[0]
    n_measurements = int(_non_na(subsampled_matrix).sum() / 2)
    total_possible = n_points * (n_points - 1) / 2
    sparsity = 1 - n_measurements / total_possible
    per_fold = n_measurements / folds
    checks, msgs = {}, []

    def check(name, ok, msg):
        checks[name] = bool(ok)
        if not ok:
            msgs.append(msg)
            if verbose:
                warnings.warn(msg)

    check("sufficient_points", n_points >= 2 * folds,
          "Very few points (%d) for %d-fold CV. Consider reducing folds or increasing subsample size." % (n_points, folds))
    check("sufficient_measurements", n_measurements >= folds * min_measurements_per_fold,
          "Insufficient measurements (%d) for %d-fold CV. Expected at least %d." % (n_measurements, folds, folds * min_measurements_per_fold))
    check("adequate_measurements_per_fold", per_fold >= min_measurements_per_fold,
          "Only ~%.1f measurements per fold (expected >= %d). Results may be unreliable." % (per_fold, min_measurements_per_fold))
    check("not_too_sparse", sparsity < 0.95,
          "Matrix is %.1f%% sparse. Such extreme sparsity may cause optimization issues." % (sparsity * 100))
    check("has_measurements", n_measurements > 0, "No measurements found in subsampled matrix!")
    return dict(all_checks_passed=all(checks.values()), checks=checks, warnings=msgs,
                diagnostics=dict(n_points=n_points, n_measurements=n_measurements, sparsity=sparsity, folds=folds,
                                 est_measurements_per_fold=per_fold, est_points_per_fold=n_points))
