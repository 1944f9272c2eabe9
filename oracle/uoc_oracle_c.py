"""ctypes wrapper of oracle/uoc_oracle_c.c (canonical-order C oracle).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libuoc_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            import sys
            sys.path.insert(0, os.path.dirname(HERE))
            from unseenobjectclustering_b200.build import build_oracle
            build_oracle()
        _lib = ctypes.CDLL(LIB)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


METRICS = {"cosine": 0, "euclidean": 1}


def select_seeds(X_planar, m, first, metric="cosine"):
    """X_planar: [d, n] float32 (row k = channel k).  Returns (selected int64 [m], seeds float32 [m, d])."""
    X = _f32(X_planar)
    d, n = X.shape
    sel = np.empty(m, dtype=np.int64)
    seeds = np.empty((m, d), dtype=np.float32)
    rc = load().uoc_oracle_select_seeds_metric(_p(X), ctypes.c_int64(n), d, ctypes.c_int64(n), m, ctypes.c_int64(int(first)),
                                               _p(sel), _p(seeds), METRICS[metric])
    assert rc == 0, rc
    return sel, seeds


def select_seeds_init(X_planar, m, init_seeds, metric="cosine"):
    """Continue from the rows of init_seeds [num_init, d].  Returns (selected int64 [m] with -1 for the given rows, seeds [m, d])."""
    X = _f32(X_planar)
    init = _f32(init_seeds)
    d, n = X.shape
    sel = np.empty(m, dtype=np.int64)
    seeds = np.empty((m, d), dtype=np.float32)
    rc = load().uoc_oracle_select_seeds_init(_p(X), ctypes.c_int64(n), d, ctypes.c_int64(n), m, _p(init), init.shape[0],
                                             _p(sel), _p(seeds), METRICS[metric])
    assert rc == 0, rc
    return sel, seeds


def label_seeds(Z, eps, metric="cosine"):
    Z = _f32(Z)
    m, d = Z.shape
    labels = np.empty(m, dtype=np.int32)
    uniq = load().uoc_oracle_label_seeds_metric(_p(Z), m, d, ctypes.c_float(eps), _p(labels), METRICS[metric])
    return labels, int(uniq)


def assign(X_planar, Z, seed_labels, num_unique, metric="cosine"):
    X = _f32(X_planar)
    Z = _f32(Z)
    d, n = X.shape
    m = Z.shape[0]
    sl = np.ascontiguousarray(seed_labels, dtype=np.int32)
    out = np.empty(n, dtype=np.int32)
    rc = load().uoc_oracle_assign_metric(_p(X), ctypes.c_int64(n), d, ctypes.c_int64(n), _p(Z), m, _p(sl), int(num_unique),
                                         _p(out), METRICS[metric])
    assert rc == 0, rc
    return out


def hill_climb(X_planar, Z, kappa, iters, metric="cosine"):
    X = _f32(X_planar)
    Zc = _f32(Z).copy()
    d, n = X.shape
    m = Zc.shape[0]
    rc = load().uoc_oracle_hill_climb_metric(_p(X), ctypes.c_int64(n), d, ctypes.c_int64(n), _p(Zc), m, ctypes.c_float(kappa),
                                             int(iters), METRICS[metric])
    assert rc == 0, rc
    return Zc


def cluster(X_planar, m, first, kappa=20.0, iters=10, eps=0.04, metric="cosine"):
    """Whole pipeline in canonical arithmetic: returns dict(selected, seeds, Z, seed_labels, num_unique, labels)."""
    sel, seeds = select_seeds(X_planar, m, first, metric)
    Z = hill_climb(X_planar, seeds, kappa, iters, metric)
    sl, uniq = label_seeds(Z, eps, metric)
    labels = assign(X_planar, Z, sl, uniq, metric)
    return dict(selected=sel, seeds=seeds, Z=Z, seed_labels=sl, num_unique=uniq, labels=labels)
