"""Host-side mirror of the reference's R entry points for the hot path.

Same names, argument meaning, defaults, validation messages and result fields as
/root/reference/R/core.R:184-528 (`euclidean_embedding`) - the R code stays the caller in a
real deployment (INTEGRATION.md shows the two-line dispatch); this module is what the parity
tests and the bench drive, and it reaches the GPU only through the C ABI (`_lib`).

An R matrix is a numpy array here: float with NaN for NA, or object/str with entries such as
3.5, "3.5", "<5", ">2", None / NaN / "NA".
"""
from __future__ import annotations

import math
import os
import warnings

import numpy as np

from . import _lib

METHOD_NAME = "b200_coloured_full_pairwise"   # the reference reports "cpp_exact_full_pairwise" (R/core.R:458)
METHOD_REPLAY = "b200_replay_full_pairwise"
METHOD_ROWBLOCK = "b200_rowblock_full_pairwise"
_MODES = {"coloured": (_lib.MODE_COLOURED, METHOD_NAME), "replay": (_lib.MODE_REPLAY, METHOD_REPLAY),
          "rowblock": (_lib.MODE_ROWBLOCK, METHOD_ROWBLOCK)}


class TopolowResult(dict):
    """The `topolow` S3 object (R/core.R:505-525) as a dict with attribute access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


# ---------------------------------------------------------------------------------------------
# Parsing (R/core.R:345-374)
# ---------------------------------------------------------------------------------------------
def _num(s):
    try:
        return float(s)
    except (TypeError, ValueError):
        return math.nan


_vnum = np.vectorize(_num, otypes=[np.float64])


def _parse_elementwise(flat):
    """Reference semantics one element at a time (kept for odd object types; the vectorised path below must agree)."""
    value = np.full(flat.shape, np.nan)
    code = np.zeros(flat.shape, dtype=np.int32)
    is_na = np.zeros(flat.shape, dtype=bool)
    for idx, x in enumerate(flat):
        if x is None:
            is_na[idx] = True
        elif isinstance(x, str):
            if x == "NA":
                is_na[idx] = True
            elif x[:1] == ">":
                code[idx] = 1
                value[idx] = _num(x[1:])
            elif x[:1] == "<":
                code[idx] = -1
                value[idx] = _num(x[1:])
            else:
                value[idx] = _num(x)
        else:
            v = _num(x)
            if math.isnan(v):
                is_na[idx] = True
            else:
                value[idx] = v
    return value, code, is_na


def _to_float(strings):
    """as.numeric() on a str array: unparsable -> NaN.  numpy's own parser when every string is a number (the usual
    case), else pandas' coercing parser, else one float() per cell."""
    if strings.size == 0:
        return np.zeros(0)
    try:
        return strings.astype(np.float64)
    except ValueError:
        pass
    try:
        import pandas as pd
        return np.array(pd.to_numeric(pd.Series(strings, dtype=object), errors="coerce"), dtype=np.float64)
    except ImportError:
        return _vnum(strings)


def parse_values(values):
    """-> (value, code, is_na) for a 1-D array of cells: numeric value (NaN where NA or unparsable), threshold code
    (+1 for '>x', -1 for '<x', 0 otherwise) and the NA mask (R/core.R:345-374: startsWith / sub / as.numeric,
    vectorised the same way - no Python-level loop over cells)."""
    flat = np.asarray(values).ravel()
    if flat.dtype.kind in "fiub":
        value = flat.astype(np.float64)
        return value, np.zeros(flat.shape, dtype=np.int32), np.isnan(value)
    if flat.dtype.kind == "O":
        # NA cells (None / NaN, usually most of the matrix) are found at C speed and skipped; only the measured cells
        # are looked at one type test each
        try:
            import pandas as pd
            is_none = np.asarray(pd.isna(flat), dtype=bool)
        except ImportError:
            is_none = np.frompyfunc(lambda x: x is None or (isinstance(x, float) and x != x), 1, 1)(flat).astype(bool)
        live = np.flatnonzero(~is_none)
        live_str = np.frompyfunc(lambda x: isinstance(x, str), 1, 1)(flat[live]).astype(bool) if live.size else np.zeros(0, bool)
        is_str = np.zeros(flat.shape, dtype=bool)
        is_str[live[live_str]] = True
        other = ~is_str & ~is_none
        value = np.full(flat.shape, np.nan)
        if other.any():
            try:
                value[other] = flat[other].astype(np.float64)
            except (TypeError, ValueError):
                return _parse_elementwise(flat)
        strs = flat[is_str].astype(str)
    else:                                   # 'U' / 'S'
        is_str = np.ones(flat.shape, dtype=bool)
        is_none = np.zeros(flat.shape, dtype=bool)
        other = ~is_str
        value = np.full(flat.shape, np.nan)
        strs = flat.astype(str)
    code = np.zeros(flat.shape, dtype=np.int32)
    is_na = is_none | (other & np.isnan(value))
    if strs.size:
        width = strs.dtype.itemsize // 4
        chars = np.ascontiguousarray(strs).view("U1").reshape(len(strs), width) if width else np.zeros((len(strs), 0), "U1")
        first = chars[:, 0] if width else np.full(len(strs), "", "U1")
        gt, lt = first == ">", first == "<"
        pref = gt | lt
        body = strs.copy()
        if pref.any() and width > 1:
            body[pref] = np.ascontiguousarray(chars[pref, 1:]).view("U%d" % (width - 1)).ravel()
        elif pref.any():
            body[pref] = ""
        sna = strs == "NA"
        sval = np.full(len(strs), np.nan)
        sval[~sna] = _to_float(body[~sna])        # the NA cells (most of a sparse matrix) are not parsed at all
        scode = np.where(gt, 1, np.where(lt, -1, 0)).astype(np.int32)
        scode[sna] = 0
        value[is_str] = sval
        code[is_str] = scode
        na_full = np.zeros(flat.shape, dtype=bool)
        na_full[is_str] = sna
        is_na |= na_full
    return value, code, is_na


def parse_dissimilarity(matrix):
    """-> (value, code, is_na): numeric value (NaN where NA or unparsable), threshold code
    (+1 for '>x', -1 for '<x', 0 otherwise) and the NA mask, for a numeric or character matrix."""
    m = np.asarray(matrix)
    value, code, is_na = parse_values(m)
    return value.reshape(m.shape), code.reshape(m.shape), is_na.reshape(m.shape)


def _is_character(matrix) -> bool:
    return np.asarray(matrix).dtype.kind not in "fiub"


def reorder_by_mean_dissimilarity(value, is_na):
    """R/core.R:269-322: ascending order of the mean off-diagonal dissimilarity (threshold
    prefixes stripped), or None when fewer than two points have a positive mean."""
    v = np.where(is_na, np.nan, value).astype(np.float64)
    np.fill_diagonal(v, np.nan)
    ok = ~np.isnan(v)
    with np.errstate(invalid="ignore", divide="ignore"):
        rs = np.where(ok, v, 0.0).sum(axis=1) / ok.sum(axis=1)
        cs = np.where(ok, v, 0.0).sum(axis=0) / ok.sum(axis=0)
    avg = (rs + cs) / 2
    avg[np.isnan(avg)] = 0
    if np.sum(avg > 0) > 1:
        return np.argsort(avg, kind="stable")
    return None


def build_problem(matrix, preserve_order=False):
    """Everything between validation and the native call (R/core.R:269-402), vectorised."""
    m = np.asarray(matrix)
    n = m.shape[0]
    value, code, is_na = parse_dissimilarity(m)
    order = None
    if n > 1 and not preserve_order:
        order = reorder_by_mean_dissimilarity(value, is_na)
        if order is not None:
            ix = np.ix_(order, order)
            m, value, code, is_na = m[ix], value[ix], code[ix], is_na[ix]
    degrees = (~is_na).sum(axis=1).astype(np.int32)          # rowSums(!is.na), diagonal included
    measured = ~is_na & ~np.isnan(value) & (value != np.inf)   # distances_numeric != Inf, NA dropped
    # which(arr.ind=TRUE) on the upper triangle walks column-major: by column j, then row i
    jj, ii = np.nonzero(np.triu(measured, 1).T)
    ei, ej = ii.astype(np.int32), jj.astype(np.int32)
    return dict(matrix=m, order=order, n=n, degrees=degrees, edge_i=ei, edge_j=ej,
                edge_dist=value[ei, ej].astype(np.float64), edge_thresh=code[ei, ej].astype(np.int32),
                value=value, code=code, is_na=is_na)


def random_initial_positions(value, code, is_na, ndim, rng):
    """R/core.R:407-415: first row 0, the others cumulative sums of U(0, 2*max/n) steps; the max
    ignores NA and threshold cells (as.numeric(as.character()) turns '<5' into NA)."""
    n = value.shape[0]
    plain = ~is_na & (code == 0) & ~np.isnan(value)
    init_step = value[plain].max() / n
    steps = rng.uniform(0.0, 2.0 * init_step, size=(n - 1, ndim))
    return np.vstack([np.zeros((1, ndim)), np.cumsum(steps, axis=0)])


def _validate(dissimilarity_matrix, ndim, mapping_max_iter, k0, cooling_rate, c_repulsion, relative_epsilon,
              convergence_counter, convergence_check_freq, initial_positions):
    """R/core.R:202-264: same checks, same messages."""
    def isnum(x):
        return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)

    if isinstance(dissimilarity_matrix, (int, np.integer)):      # the sparse entry: only the point count exists
        n_rows = int(dissimilarity_matrix)
    else:
        if not isinstance(dissimilarity_matrix, np.ndarray) or dissimilarity_matrix.ndim != 2:
            raise ValueError("dissimilarity_matrix must be a matrix")
        if dissimilarity_matrix.shape[0] != dissimilarity_matrix.shape[1]:
            raise ValueError("dissimilarity_matrix must be square")
        n_rows = dissimilarity_matrix.shape[0]
    if not isnum(ndim) or ndim < 1 or ndim != round(ndim):
        raise ValueError("ndim must be a positive integer")
    if not isnum(mapping_max_iter) or mapping_max_iter < 1 or mapping_max_iter != round(mapping_max_iter):
        raise ValueError("mapping_max_iter must be a positive integer")
    if not isnum(k0) or k0 <= 0:
        raise ValueError("k0 must be a positive number")
    if k0 > 30:
        warnings.warn("High k0 value (> 30) may lead to instability")
    if not isnum(cooling_rate) or cooling_rate <= 0 or cooling_rate >= 1:
        raise ValueError("cooling_rate must be between 0 and 1")
    if not isnum(c_repulsion) or c_repulsion <= 0:
        raise ValueError("c_repulsion must be a positive number")
    if not isnum(relative_epsilon) or relative_epsilon <= 0:
        raise ValueError("relative_epsilon must be a positive number")
    if not isnum(convergence_counter) or convergence_counter < 1 or convergence_counter != round(convergence_counter):
        raise ValueError("convergence_counter must be a positive integer")
    if not isnum(convergence_check_freq) or convergence_check_freq < 1:
        raise ValueError("convergence_check_freq must be a positive integer")
    if initial_positions is not None:
        if not isinstance(initial_positions, np.ndarray) or initial_positions.ndim != 2:
            raise ValueError("initial_positions must be a matrix")
        if initial_positions.shape[0] != n_rows:
            raise ValueError("initial_positions must have same number of rows as dissimilarity_matrix")
        if initial_positions.shape[1] != ndim:
            raise ValueError("initial_positions must have ndim columns")
    if n_rows < 2:
        raise ValueError("dissimilarity_matrix must have at least 2 rows/columns")


def euclidean_embedding(dissimilarity_matrix, ndim, mapping_max_iter=1000, k0=None, cooling_rate=None,
                        c_repulsion=None, relative_epsilon=1e-4, convergence_counter=5, initial_positions=None,
                        write_positions_to_csv=False, output_dir=None, verbose=False, convergence_check_freq=3,
                        preserve_order=False, *, rownames=None, mode="coloured", precision="f32", seed=0,
                        pair_order=None, device=0, rng=None, return_est_distances=True):
    """Drop-in for the reference's euclidean_embedding() (R/core.R:184-528).

    Keyword-only arguments after `preserve_order` are B200 extensions: `mode` ("coloured" production
    schedule or "replay" of a seeded std::mt19937 pair permutation), `precision` ("f32" | "f64"),
    `seed`, `pair_order` (replay), `device`.  `rownames` stands in for rownames(dissimilarity_matrix).
    Result fields: positions, est_distances, mae, iter, parameters, convergence (+ rownames, order).
    """
    if k0 is None or cooling_rate is None or c_repulsion is None:
        raise TypeError("k0, cooling_rate and c_repulsion are required")
    _validate(dissimilarity_matrix, ndim, mapping_max_iter, k0, cooling_rate, c_repulsion, relative_epsilon,
              convergence_counter, convergence_check_freq, initial_positions)
    value0, code0, na0 = parse_dissimilarity(dissimilarity_matrix)
    finite = value0[~na0 & np.isfinite(value0)]
    if np.sum(finite != 0) == 0:
        warnings.warn("No finite non-zero dissimilarities found. Results may be unreliable.")

    prob = build_problem(dissimilarity_matrix, preserve_order)
    n = prob["n"]
    order = prob["order"]
    names = None if rownames is None else list(rownames)
    if order is not None and names is not None:
        names = [names[i] for i in order]
    if initial_positions is not None and order is not None and rownames is not None:
        # R/core.R:325-333 re-aligns by row names only when both sides have them
        initial_positions = np.asarray(initial_positions)[order]
    if len(prob["edge_i"]) == 0:
        raise ValueError("No valid off-diagonal measurements found in dissimilarity matrix")
    if initial_positions is None:
        initial_positions = random_initial_positions(prob["value"], prob["code"], prob["is_na"], int(ndim),
                                                     rng or np.random.default_rng(seed))

    mode_c = _MODES[mode][0]
    prec_c = {"f32": _lib.PREC_F32, "f64": _lib.PREC_F64_EXACT}[precision]
    res = _lib.fit(initial_positions, prob["degrees"], prob["edge_i"], prob["edge_j"], prob["edge_dist"],
                   prob["edge_thresh"], int(mapping_max_iter), k0, cooling_rate, c_repulsion, relative_epsilon,
                   int(convergence_counter), int(convergence_check_freq), verbose=verbose, mode=mode_c,
                   precision=prec_c, seed=seed, pair_order=pair_order, device=device)
    positions = res["positions"]

    est = None
    mae = math.nan
    if return_est_distances:
        est = _lib.est_distances(positions, device)                  # R/core.R:474
        # R/core.R:479-481: as.numeric() of the (reordered) input: threshold strings are NA, the
        # diagonal and both triangles of a numeric matrix count
        raw_ok = ~prob["is_na"] & (prob["code"] == 0) & ~np.isnan(prob["value"])
        if raw_ok.any():
            mae = float(np.mean(np.abs(prob["value"][raw_ok] - est[raw_ok])))

    if write_positions_to_csv:
        if output_dir is None:
            raise ValueError("An 'output_dir' must be provided when 'write_positions_to_csv' is TRUE.")
        os.makedirs(output_dir, exist_ok=True)
        fn = "Positions_dim_%d_k0_%.4f_cooling_%.4f_c_repulsion_%.4f.csv" % (ndim, k0, cooling_rate, c_repulsion)
        rn = names or [str(i + 1) for i in range(n)]
        with open(os.path.join(output_dir, fn), "w") as f:
            f.write('"",' + ",".join('"V%d"' % (i + 1) for i in range(int(ndim))) + "\n")
            for r, row in zip(rn, positions):
                f.write('"%s",' % r + ",".join(repr(float(x)) for x in row) + "\n")

    return TopolowResult(
        positions=positions, est_distances=est, mae=mae, iter=res["iterations"],
        parameters=dict(ndim=ndim, k0=k0, cooling_rate=cooling_rate, c_repulsion=c_repulsion,
                        method=_MODES[mode][1]),
        convergence=dict(achieved=res["converged"], error=res["final_mae"], final_k=res["final_k"]),
        rownames=names, order=order, native=res)


# ---------------------------------------------------------------------------------------------
# Sparse entry: the same fit from a (row, column, value) table - nothing n x n is ever built
# ---------------------------------------------------------------------------------------------
def build_problem_coo(n, rows, cols, values, preserve_order=False, diagonal=True):
    """What build_problem() derives from the dense matrix (R/core.R:269-402), from its non-NA off-diagonal cells.

    `rows`, `cols` are 0-based; `values` numbers or strings ('<x', '>x', 'NA' is dropped).  A pair may be listed in one
    or both orientations (a symmetric matrix written out in full lists every pair twice); the first listing wins.
    `diagonal`: the matrix this table stands for has its diagonal filled in (the package's converters write 0 there),
    so rowSums(!is.na) counts one more cell per point (R/core.R:340-341).
    Identical to the dense path: same degrees, same order, same edges in which(arr.ind = TRUE) order."""
    rows = np.asarray(rows, dtype=np.int64).ravel()
    cols = np.asarray(cols, dtype=np.int64).ravel()
    if len(rows) != len(cols) or len(rows) != np.size(values):
        raise ValueError("rows, cols and values must have the same length")
    if len(rows) and (rows.min() < 0 or cols.min() < 0 or rows.max() >= n or cols.max() >= n):
        raise ValueError("row / column index out of range")
    value, code, is_na = parse_values(values)
    keep = ~is_na & (rows != cols)
    lo, hi = np.minimum(rows, cols)[keep], np.maximum(rows, cols)[keep]
    value, code = value[keep], code[keep]
    key = hi * np.int64(n) + lo                         # column-major position of the upper-triangle cell
    key, first = np.unique(key, return_index=True)      # sorted by (column, row) = which(arr.ind = TRUE) order; first listing wins
    lo, hi, value, code = lo[first], hi[first], value[first], code[first]
    degrees = (np.bincount(lo, minlength=n) + np.bincount(hi, minlength=n) + (1 if diagonal else 0)).astype(np.int32)
    order = None
    if n > 1 and not preserve_order:
        # R/core.R:269-322 on the symmetric matrix: row mean = column mean = mean over the point's non-NA, non-diagonal
        # cells whose value parses (threshold prefixes stripped)
        ok = ~np.isnan(value)
        tot = np.bincount(lo[ok], weights=value[ok], minlength=n) + np.bincount(hi[ok], weights=value[ok], minlength=n)
        cnt = np.bincount(lo[ok], minlength=n) + np.bincount(hi[ok], minlength=n)
        with np.errstate(invalid="ignore", divide="ignore"):
            avg = tot / cnt
        avg[np.isnan(avg)] = 0
        if np.sum(avg > 0) > 1:
            order = np.argsort(avg, kind="stable")
            rank = np.empty(n, dtype=np.int64)
            rank[order] = np.arange(n)
            a, b = rank[lo], rank[hi]
            lo, hi = np.minimum(a, b), np.maximum(a, b)
            resort = np.argsort(hi * np.int64(n) + lo, kind="stable")
            lo, hi, value, code = lo[resort], hi[resort], value[resort], code[resort]
            degrees = degrees[order]
    measured = ~np.isnan(value) & (value != np.inf)
    plain = (code == 0) & ~np.isnan(value)
    return dict(n=int(n), order=order, degrees=degrees, edge_i=lo[measured].astype(np.int32), edge_j=hi[measured].astype(np.int32),
                edge_dist=value[measured].astype(np.float64), edge_thresh=code[measured].astype(np.int32),
                cell_i=lo, cell_j=hi, cell_value=value, cell_code=code, plain=plain, diagonal=diagonal)


def euclidean_embedding_coo(n, rows, cols, values, ndim, mapping_max_iter=1000, k0=None, cooling_rate=None, c_repulsion=None,
                            relative_epsilon=1e-4, convergence_counter=5, initial_positions=None, verbose=False,
                            convergence_check_freq=3, preserve_order=False, *, rownames=None, diagonal=True, mode="coloured",
                            precision="f32", seed=0, device=0, rng=None, extra_pairs=None):
    """euclidean_embedding() (R/core.R:184-528) for inputs too large to hold as an n x n matrix: the dissimilarities
    come as a table, the result carries `est_distances` only for the measured pairs (`pairs`, in the problem's own
    order) and for `extra_pairs` = (i, j) arrays of further cells (held-out ones, say) - `est_extra`.  Positions,
    convergence fields and `mae` are those of the dense call on the matrix the table stands for (R/core.R:479-481
    averages |input - estimate| over every numeric cell of the symmetric matrix: both orientations of a pair and, when
    `diagonal`, the n zeros of the diagonal)."""
    if k0 is None or cooling_rate is None or c_repulsion is None:
        raise TypeError("k0, cooling_rate and c_repulsion are required")
    _validate(int(n), ndim, mapping_max_iter, k0, cooling_rate, c_repulsion, relative_epsilon, convergence_counter,
              convergence_check_freq, initial_positions)
    prob = build_problem_coo(n, rows, cols, values, preserve_order, diagonal)
    if len(prob["edge_i"]) == 0:
        raise ValueError("No valid off-diagonal measurements found in dissimilarity matrix")
    fin = prob["cell_value"][np.isfinite(prob["cell_value"])]
    if np.sum(fin != 0) == 0:
        warnings.warn("No finite non-zero dissimilarities found. Results may be unreliable.")
    order = prob["order"]
    names = None if rownames is None else list(rownames)
    if order is not None and names is not None:
        names = [names[i] for i in order]
    if initial_positions is not None and order is not None and rownames is not None:
        initial_positions = np.asarray(initial_positions)[order]
    if initial_positions is None:
        vmax = prob["cell_value"][prob["plain"]].max() if prob["plain"].any() else 0.0
        if diagonal:
            vmax = max(vmax, 0.0)
        steps = (rng or np.random.default_rng(seed)).uniform(0.0, 2.0 * vmax / n, size=(n - 1, int(ndim)))     # R/core.R:407-415
        initial_positions = np.vstack([np.zeros((1, int(ndim))), np.cumsum(steps, axis=0)])
    res = _lib.fit(initial_positions, prob["degrees"], prob["edge_i"], prob["edge_j"], prob["edge_dist"], prob["edge_thresh"],
                   int(mapping_max_iter), k0, cooling_rate, c_repulsion, relative_epsilon, int(convergence_counter),
                   int(convergence_check_freq), verbose=verbose, mode=_MODES[mode][0],
                   precision={"f32": _lib.PREC_F32, "f64": _lib.PREC_F64_EXACT}[precision], seed=seed, device=device)
    positions = res["positions"]
    ci, cj = prob["cell_i"], prob["cell_j"]
    est = np.linalg.norm(positions[ci] - positions[cj], axis=1)
    pl = prob["plain"]
    cells = 2 * int(pl.sum()) + (n if diagonal else 0)
    mae = float(2.0 * np.abs(prob["cell_value"][pl] - est[pl]).sum() / cells) if cells else math.nan
    est_extra = None
    if extra_pairs is not None:
        xi, xj = (np.asarray(a, dtype=np.int64) for a in extra_pairs)
        if order is not None:                            # the caller's numbering -> the problem's
            rank = np.empty(n, dtype=np.int64)
            rank[order] = np.arange(n)
            xi, xj = rank[xi], rank[xj]
        est_extra = np.linalg.norm(positions[xi] - positions[xj], axis=1)
    return TopolowResult(
        positions=positions, est_distances=est, pairs=(ci, cj), est_extra=est_extra, mae=mae, iter=res["iterations"],
        parameters=dict(ndim=ndim, k0=k0, cooling_rate=cooling_rate, c_repulsion=c_repulsion, method=_MODES[mode][1]),
        convergence=dict(achieved=res["converged"], error=res["final_mae"], final_k=res["final_k"]),
        rownames=names, order=order, native=res)
