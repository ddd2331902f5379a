"""ctypes access to the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; nothing under topolow_b200/ does.

  optimize_layout_exact(...)      -> oracle/topolow_oracle.cpp (restatement of
                                     /root/reference/src/optimization.cpp:108-382)
  ref_optimize_layout_exact(...)  -> oracle/_ref/libtopolow_ref.so: the reference's own
                                     source compiled against oracle/ref_shim (when built)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def build(ref: bool = True) -> None:
    """Compile liboracle*.so and, when /root/reference exists, _ref/libtopolow_ref.so."""
    targets = ["all"]
    if ref and os.path.exists("/root/reference/src/optimization.cpp"):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def _lib(kind: str = "O2") -> C.CDLL:
    if kind in _LIBS:
        return _LIBS[kind]
    name = {"O2": "liboracle.so", "fast": "liboracle_fast.so"}[kind]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build(ref=False)
    lib = C.CDLL(path)
    lib.oracle_optimize_layout_exact.restype = C.c_int
    lib.oracle_optimize_layout_exact.argtypes = [
        C.c_int64, C.c_int, _dp, _dp, _ip, _ip, C.c_int64, _ip, _ip, _dp, _ip, C.c_int,
        C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_uint32,
        _i32p, C.c_int64, _dp, _ip, _ip, _dp, _dp, _dp, _i64p]
    lib.oracle_pair_orders.restype = C.c_int
    lib.oracle_pair_orders.argtypes = [C.c_int64, C.c_int, C.c_uint32, _i32p]
    lib.oracle_edge_error.restype = None
    lib.oracle_edge_error.argtypes = [_dp, C.c_int64, C.c_int, C.c_int64, _ip, _ip, _dp, _ip, _dp, _i64p]
    _LIBS[kind] = lib
    return lib


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libtopolow_ref.so"))


def _ref_lib() -> C.CDLL:
    if "ref" in _LIBS:
        return _LIBS["ref"]
    lib = C.CDLL(os.path.join(_HERE, "_ref", "libtopolow_ref.so"))
    lib.ref_optimize_layout_exact.restype = C.c_int
    lib.ref_optimize_layout_exact.argtypes = [
        C.c_int, C.c_int, _dp, _dp, _ip, _ip, C.c_int, _ip, _ip, _dp, _ip, C.c_int, C.c_double,
        C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_uint, _dp, _ip, _ip, _dp, _dp,
        C.c_char_p, C.c_int]
    _LIBS["ref"] = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class OracleError(RuntimeError):
    pass


def dense_from_edges(n, edge_i, edge_j, edge_dist, edge_thresh):
    """Dense symmetric Inf-filled matrices as R/core.R:429-436 hands them over
    (column-major; symmetric, so the memory image equals the row-major one)."""
    dm = np.full((n, n), np.inf)
    tm = np.zeros((n, n), dtype=np.int32)
    dm[edge_i, edge_j] = edge_dist
    dm[edge_j, edge_i] = edge_dist
    tm[edge_i, edge_j] = edge_thresh
    tm[edge_j, edge_i] = edge_thresh
    return dm, tm


def optimize_layout_exact(initial_positions, degrees, edge_i, edge_j, edge_dist, edge_thresh, n_iter,
                          k0, cooling_rate, c_repulsion, relative_epsilon=1e-4, convergence_window=5,
                          convergence_check_freq=3, *, seed=0, pair_order=None, sample_pairs=None,
                          dense=True, dissimilarity_matrix=None, threshold_matrix=None, kind="O2",
                          trace=False):
    """Run the restated loop.  `initial_positions` is (n, dim), any layout; the
    result dict mirrors src/optimization.cpp:375-381 plus `visited` (pair visits)."""
    lib = _lib(kind)
    init = np.asarray(initial_positions, dtype=np.float64)
    n, dim = init.shape
    init_cm = np.asfortranarray(init)  # column-major like an R matrix
    ei, ej = _i32(edge_i), _i32(edge_j)
    ed, et = _f64(edge_dist), _i32(edge_thresh)
    deg = _i32(degrees)
    dm = tm = None
    if dense:
        if dissimilarity_matrix is None:
            dm, tm = dense_from_edges(n, ei, ej, ed, et)
        else:
            dm = np.asfortranarray(dissimilarity_matrix, dtype=np.float64)
            tm = np.asfortranarray(threshold_matrix, dtype=np.int32)
    order_mode, po, ppi = 0, None, 0
    if pair_order is not None:
        po = np.ascontiguousarray(pair_order, dtype=np.int32)
        assert po.ndim == 3 and po.shape[0] >= n_iter and po.shape[2] == 2
        order_mode, ppi = 1, po.shape[1]
    elif sample_pairs is not None:
        order_mode, ppi = 2, int(sample_pairs)
    out = np.empty((n, dim), dtype=np.float64, order="F")
    conv, iters = C.c_int(0), C.c_int(0)
    fmae, fk = C.c_double(0), C.c_double(0)
    visited = C.c_int64(0)
    tr = np.full(n_iter, np.nan) if trace else None
    rc = lib.oracle_optimize_layout_exact(
        n, dim, init_cm.ctypes.data_as(_dp), _p(dm, _dp), _p(tm, _ip), deg.ctypes.data_as(_ip), len(ei),
        ei.ctypes.data_as(_ip), ej.ctypes.data_as(_ip), ed.ctypes.data_as(_dp), et.ctypes.data_as(_ip),
        int(n_iter), float(k0), float(cooling_rate), float(c_repulsion), float(relative_epsilon),
        int(convergence_window), int(convergence_check_freq), order_mode, int(seed) & 0xFFFFFFFF,
        _p(po, _i32p), ppi, out.ctypes.data_as(_dp), C.byref(conv), C.byref(iters), C.byref(fmae),
        C.byref(fk), _p(tr, _dp), C.byref(visited))
    if rc == 1:
        raise OracleError("Need at least 2 points for embedding")
    if rc == 2:
        raise OracleError("Numerical instability at iteration %d. Reduce k0 or c_repulsion." % iters.value)
    res = dict(positions=np.ascontiguousarray(out), converged=bool(conv.value), iterations=iters.value,
               final_mae=fmae.value, final_k=fk.value, visited=visited.value)
    if trace:
        res["trace_mae"] = tr
    return res


def ref_optimize_layout_exact(initial_positions, degrees, edge_i, edge_j, edge_dist, edge_thresh, n_iter,
                              k0, cooling_rate, c_repulsion, relative_epsilon=1e-4, convergence_window=5,
                              convergence_check_freq=3, *, seed=0):
    """The reference's own optimize_layout_exact_cpp (compiled from
    /root/reference/src/optimization.cpp) with std::random_device forced to `seed`."""
    lib = _ref_lib()
    init = np.asarray(initial_positions, dtype=np.float64)
    n, dim = init.shape
    init_cm = np.asfortranarray(init).copy(order="F")
    ei, ej = _i32(edge_i), _i32(edge_j)
    ed, et = _f64(edge_dist), _i32(edge_thresh)
    deg = _i32(degrees)
    dm, tm = dense_from_edges(n, ei, ej, ed, et)
    out = np.empty((n, dim), dtype=np.float64, order="F")
    conv, iters = C.c_int(0), C.c_int(0)
    fmae, fk = C.c_double(0), C.c_double(0)
    err = C.create_string_buffer(512)
    rc = lib.ref_optimize_layout_exact(
        n, dim, init_cm.ctypes.data_as(_dp), dm.ctypes.data_as(_dp), tm.ctypes.data_as(_ip),
        deg.ctypes.data_as(_ip), len(ei), ei.ctypes.data_as(_ip), ej.ctypes.data_as(_ip),
        ed.ctypes.data_as(_dp), et.ctypes.data_as(_ip), int(n_iter), float(k0), float(cooling_rate),
        float(c_repulsion), float(relative_epsilon), int(convergence_window), int(convergence_check_freq),
        int(seed) & 0xFFFFFFFF, out.ctypes.data_as(_dp), C.byref(conv), C.byref(iters), C.byref(fmae),
        C.byref(fk), err, 512)
    if rc != 0:
        raise OracleError(err.value.decode())
    return dict(positions=np.ascontiguousarray(out), converged=bool(conv.value), iterations=iters.value,
                final_mae=fmae.value, final_k=fk.value)


def pair_orders(n: int, n_iter: int, seed: int) -> np.ndarray:
    """(n_iter, P, 2) int32: the pair order std::shuffle(mt19937(seed)) draws per iteration."""
    P = n * (n - 1) // 2
    out = np.empty((n_iter, P, 2), dtype=np.int32)
    _lib().oracle_pair_orders(n, n_iter, int(seed) & 0xFFFFFFFF, out.ctypes.data_as(_i32p))
    return out


def edge_error(positions, edge_i, edge_j, edge_dist, edge_thresh):
    pos = np.asfortranarray(positions, dtype=np.float64)
    n, dim = pos.shape
    ei, ej, ed, et = _i32(edge_i), _i32(edge_j), _f64(edge_dist), _i32(edge_thresh)
    tot, cnt = C.c_double(0), C.c_int64(0)
    _lib().oracle_edge_error(pos.ctypes.data_as(_dp), n, dim, len(ei), ei.ctypes.data_as(_ip),
                             ej.ctypes.data_as(_ip), ed.ctypes.data_as(_dp), et.ctypes.data_as(_ip),
                             C.byref(tot), C.byref(cnt))
    return tot.value, cnt.value


def _relaxed_lib() -> C.CDLL:
    if "relaxed" in _LIBS:
        return _LIBS["relaxed"]
    path = os.path.join(_HERE, "librelaxed.so")
    if not os.path.exists(path):
        build(ref=False)
    lib = C.CDLL(path)
    lib.relaxed_optimize_layout.restype = C.c_int
    lib.relaxed_optimize_layout.argtypes = [
        C.c_int64, C.c_int, _dp, _ip, C.c_int64, _ip, _ip, _dp, _ip, C.c_int, C.c_double, C.c_double, C.c_double,
        C.c_double, C.c_int, C.c_int, _i32p, C.c_int, C.c_uint64, _dp, _ip, _ip, _dp, _dp, _dp, _ip]
    _LIBS["relaxed"] = lib
    return lib


def relaxed_optimize_layout(initial_positions, degrees, edge_i, edge_j, edge_dist, edge_thresh, n_iter, k0,
                            cooling_rate, c_repulsion, relative_epsilon=1e-4, convergence_window=5,
                            convergence_check_freq=3, *, seed=0, slot_of_point=None, rotate=True, trace=False):
    """oracle/relaxed_oracle.cpp: the row-block scheme of topolow_b200/csrc/rowblock.cu in FP64.
    slot_of_point (optional): the relabelling the GPU uses (topolow_b200.rowblock.slot_order(n))."""
    lib = _relaxed_lib()
    init = np.asfortranarray(np.asarray(initial_positions, dtype=np.float64))
    n, dim = init.shape
    ei, ej, ed, et, deg = _i32(edge_i), _i32(edge_j), _f64(edge_dist), _i32(edge_thresh), _i32(degrees)
    pos = None
    if slot_of_point is not None:
        sop = np.asarray(slot_of_point, dtype=np.int64)
        pos = np.empty(n, dtype=np.int32)
        pos[sop] = np.arange(n, dtype=np.int32)          # point_of_slot
    out = np.empty((n, dim), dtype=np.float64, order="F")
    conv, iters, run = C.c_int(0), C.c_int(0), C.c_int(0)
    fmae, fk = C.c_double(0), C.c_double(0)
    tr = np.full(max(int(n_iter), 1), np.nan) if trace else None
    rc = lib.relaxed_optimize_layout(
        n, dim, init.ctypes.data_as(_dp), deg.ctypes.data_as(_ip), len(ei), ei.ctypes.data_as(_ip),
        ej.ctypes.data_as(_ip), ed.ctypes.data_as(_dp), et.ctypes.data_as(_ip), int(n_iter), float(k0),
        float(cooling_rate), float(c_repulsion), float(relative_epsilon), int(convergence_window),
        int(convergence_check_freq), _p(pos, _i32p), int(bool(rotate)), int(seed) & 0xFFFFFFFFFFFFFFFF,
        out.ctypes.data_as(_dp), C.byref(conv), C.byref(iters), C.byref(fmae), C.byref(fk), _p(tr, _dp), C.byref(run))
    if rc == 1:
        raise OracleError("Need at least 2 points for embedding")
    if rc == 2:
        raise OracleError("Numerical instability at iteration %d. Reduce k0 or c_repulsion." % iters.value)
    res = dict(positions=np.ascontiguousarray(out), converged=bool(conv.value), iterations=iters.value,
               final_mae=fmae.value, final_k=fk.value, iterations_run=run.value)
    if trace:
        res["trace_mae"] = tr
    return res
