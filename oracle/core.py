"""ctypes access to ``liboracle_core.so`` (the C restatement; TEST INFRASTRUCTURE / CPU BASELINE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_core.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_core.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_core.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.orc_rand_u01_f32.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_size_t]
        L.orc_rand_u01_f64.argtypes = [f64p, C.c_uint64, C.c_uint64, C.c_size_t]
        L.orc_rand_ksum.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t]
        L.orc_rand_ksum.restype = C.c_uint64
        L.orc_ksum_f32.argtypes = [f32p, C.c_size_t]
        L.orc_ksum_f32.restype = C.c_uint64
        L.orc_affine_f32.argtypes = [f32p, f32p, C.c_float, C.c_float, C.c_size_t]
        L.orc_affine_f64.argtypes = [f64p, f64p, C.c_double, C.c_double, C.c_size_t]
        L.orc_sum_f32.argtypes = [f32p, C.c_size_t, C.c_int, C.c_int]
        L.orc_sum_f32.restype = C.c_float
        L.orc_sum_f64.argtypes = [f64p, C.c_size_t, C.c_int, C.c_int]
        L.orc_sum_f64.restype = C.c_double
        L.orc_fold_sum_f32.argtypes = [f32p, C.c_size_t]
        L.orc_fold_sum_f32.restype = C.c_float
        L.orc_sumdim_f32.argtypes = [f32p, C.c_size_t, C.c_size_t, C.c_size_t, f32p, C.c_int, C.c_int]
        L.orc_sumdim_f64.argtypes = [f64p, C.c_size_t, C.c_size_t, C.c_size_t, f64p, C.c_int, C.c_int]
        L.orc_max_f32.argtypes = [f32p, C.c_size_t, f32p]
        L.orc_min_f32.argtypes = [f32p, C.c_size_t, f32p]
        L.orc_workers_run.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_uint64, C.c_float, C.c_float, C.c_int,
                                      C.c_int, f32p, f64p]
        L.orc_workers_run.restype = C.c_double
        L.orc_num_procs.restype = C.c_int
        L.orc_num_usable_procs.restype = C.c_int
        L.orc_rand_stats_mt.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def rand_u01_f32(seed: int, start: int, n: int) -> np.ndarray:
    x = np.empty(n, dtype=np.float32)
    lib().orc_rand_u01_f32(_p(x, C.c_float), seed, start, n)
    return x


def rand_ksum(seed: int, start: int, n: int) -> int:
    return int(lib().orc_rand_ksum(seed, start, n))


def ksum_f32(x: np.ndarray) -> int:
    x = np.ascontiguousarray(x, dtype=np.float32)
    return int(lib().orc_ksum_f32(_p(x, C.c_float), x.size))


def affine_f32(x: np.ndarray, a: float, b: float) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_affine_f32(_p(y, C.c_float), _p(x, C.c_float), a, b, x.size)
    return y


def sum_f32(x: np.ndarray, lanes: int = 8, inter: int = 4) -> np.float32:
    x = np.ascontiguousarray(x, dtype=np.float32)
    return np.float32(lib().orc_sum_f32(_p(x, C.c_float), x.size, lanes, inter))


def sum_f64(x: np.ndarray, lanes: int = 4, inter: int = 4) -> np.float64:
    x = np.ascontiguousarray(x, dtype=np.float64)
    return np.float64(lib().orc_sum_f64(_p(x, C.c_double), x.size, lanes, inter))


def sumdim_f32(x: np.ndarray, inner: int, red: int, outer: int, lanes: int = 8, inter: int = 4) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(inner * outer, dtype=np.float32)
    lib().orc_sumdim_f32(_p(x, C.c_float), inner, red, outer, _p(out, C.c_float), lanes, inter)
    return out


def max_f32(x: np.ndarray) -> np.float32:
    x = np.ascontiguousarray(x, dtype=np.float32)
    r = C.c_float()
    if lib().orc_max_f32(_p(x, C.c_float), x.size, C.byref(r)) != 0:
        raise ValueError("reducing over an empty collection is not allowed")
    return np.float32(r.value)


def min_f32(x: np.ndarray) -> np.float32:
    x = np.ascontiguousarray(x, dtype=np.float32)
    r = C.c_float()
    if lib().orc_min_f32(_p(x, C.c_float), x.size, C.byref(r)) != 0:
        raise ValueError("reducing over an empty collection is not allowed")
    return np.float32(r.value)


def workers_run(op: int, nworkers: int, n_per: int, seed: int, a: float, b: float, warm: int, iters: int):
    """P single-threaded workers, one chunk each.  Returns (best_s, mean_s, result)."""
    r = C.c_float()
    mean = C.c_double()
    best = lib().orc_workers_run(op, nworkers, n_per, seed, a, b, warm, iters, C.byref(r), C.byref(mean))
    return float(best), float(mean.value), np.float32(r.value)


def num_procs() -> int:
    return int(lib().orc_num_procs())


def num_usable_procs() -> int:
    return int(lib().orc_num_usable_procs())


def rand_stats(seed: int, start: int, n: int, a: float, b: float, yscale: int = 25, nthreads: int = 0) -> dict:
    """Exact statistics of x = rand_u01(seed, start..start+n) and y = fl(fl(a*x)+b) (multi-threaded, integer accumulation):
    ``ksum`` (sum x = ksum * 2^-24), ``ysum`` (sum y = ysum * 2^-yscale, valid when ``inexact == 0``), ``ymax``, ``xmax``."""
    out = (C.c_uint64 * 5)()
    lib().orc_rand_stats_mt(seed, start, n, a, b, yscale, nthreads or num_usable_procs(), out)
    return {"ksum": int(out[0]), "ysum": int(out[1]), "yscale": yscale, "inexact": int(out[3]),
            "ymax": np.array([out[2]], dtype=np.uint32).view(np.float32)[0], "xmax": np.array([out[4]], dtype=np.uint32).view(np.float32)[0]}
