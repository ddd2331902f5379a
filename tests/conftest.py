import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CUDA library and the oracle exist (both travel pre-built to the GPU box)."""
    from oracle import cpu_oracle
    from topolow_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        cpu_oracle.build(ref=True)
    yield


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))
    p = z["params"]
    args = (z["initial_positions"], z["degrees"], z["edge_i"], z["edge_j"], z["edge_dist"], z["edge_thresh"],
            int(p[0]), float(p[1]), float(p[2]), float(p[3]), float(p[4]), int(p[5]), int(p[6]))
    return z, args, int(p[7])


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, "fixtures", name + ".npz"))
    return {k: z[k] for k in z.files}


def small_problem(n, d, density, seed, thresholds=True):
    """Random low-rank problem in edge-list form (edges in R's which(arr.ind) order)."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)) * 3
    iu = np.triu_indices(n, 1)
    keep = rng.random(len(iu[0])) < density
    keep[rng.integers(0, len(keep))] = True
    ei, ej = iu[0][keep], iu[1][keep]
    order = np.lexsort((ei, ej))
    ei, ej = ei[order].astype(np.int32), ej[order].astype(np.int32)
    ed = np.linalg.norm(X[ei] - X[ej], axis=1) * (1 + 0.05 * rng.normal(size=len(ei)))
    ed = np.maximum(ed, 0.1)
    et = (rng.choice([0, 0, 0, 0, 1, -1], size=len(ei)) if thresholds else np.zeros(len(ei))).astype(np.int32)
    deg = (np.bincount(ei, minlength=n) + np.bincount(ej, minlength=n) + 1).astype(np.int32)
    init = np.vstack([np.zeros((1, d)), np.cumsum(rng.uniform(0, 2 * ed.max() / n, size=(n - 1, d)), axis=0)])
    return init, deg, ei, ej, ed, et


def random_r_matrix(n, density, seed, thresholds=True):
    """An R-style character matrix: NaN = NA, strings '<x' / '>x', zero diagonal, symmetric."""
    rng = np.random.default_rng(seed)
    m = np.full((n, n), None, dtype=object)
    for i in range(n):
        m[i, i] = 0.0
        for j in range(i + 1, n):
            if rng.random() < density:
                v = round(float(rng.uniform(0.5, 8.0)), 3)
                u = rng.random()
                s = v if (not thresholds or u > 0.2) else ((">" if u < 0.1 else "<") + repr(v))
                m[i, j] = s
                m[j, i] = s
    return m


# ---- the .Call shim (integration/r_shim.c) built against the stand-in R headers of integration/r_stub/ ----
def build_r_shim_harness(out_path, real_library):
    """Compile shim + stand-in runtime + harness; linked with the recording fake (CPU) or with libtopolow_b200.so."""
    import subprocess
    stub = os.path.join(ROOT, "integration", "r_stub")
    cmd = ["gcc", "-std=gnu11", "-O1", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
           "-I", stub, "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "integration", "r_shim.c"),
           os.path.join(stub, "r_stub.c"), os.path.join(stub, "harness.c"), "-o", out_path]
    if real_library:
        libdir = os.path.join(ROOT, "topolow_b200", "lib")
        cmd += ["-L", libdir, "-ltopolow_b200", "-Wl,-rpath," + libdir]
    else:
        cmd.insert(-2, os.path.join(stub, "fake_topolow.c"))
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out_path


def write_problem_bin(path, init, deg, ei, ej, ed, et, holdout=None):
    hi, hj, ht = holdout if holdout is not None else (np.zeros(0, np.int32),) * 2 + (np.zeros(0),)
    with open(path, "wb") as f:
        np.array([init.shape[0], init.shape[1], len(ei), len(hi)], dtype=np.int64).tofile(f)
        np.asfortranarray(init, dtype=np.float64).T.tofile(f)            # column-major, like an R matrix
        for a, t in ((deg, np.int32), (ei, np.int32), (ej, np.int32), (ed, np.float64), (et, np.int32), (hi, np.int32),
                     (hj, np.int32), (ht, np.float64)):
            np.ascontiguousarray(a, dtype=t).tofile(f)
