"""ctypes loader for ``oracle/libpmm_oracle.so`` -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module (see ``pmm_oracle.c``).  The
arithmetic lives in the C file; this wrapper only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpmm_oracle.so")
    src = os.path.join(_HERE, "pmm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libpmm_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        tree = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_az_partition.argtypes = tree + [_i32p]
        L.oracle_az_partition.restype = None
        L.oracle_inside.argtypes = tree + [_i32p, _f64p, _f64p, C.c_double, C.c_double, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_inside.restype = C.c_double
        L.oracle_estep.argtypes = tree + [_i32p, _f64p, _f64p, C.c_double, C.c_double, C.c_int, C.c_int,
                                          C.c_int, _f64p, _f64p, C.c_void_p, C.c_void_p]
        L.oracle_estep.restype = C.c_double
        L.oracle_mstep_stats.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, _u8p, C.c_double, C.c_double,
                                         C.c_double, _f64p]
        L.oracle_mstep_stats.restype = C.c_double
        L.oracle_mstep_brlen.argtypes = [C.c_int, _f64p, _u8p, C.c_double, C.c_double, C.c_double,
                                         C.c_double, _f64p]
        L.oracle_mstep_brlen.restype = None
        L.oracle_mstep_nu.argtypes = [C.c_int, _f64p, _u8p, _f64p]
        L.oracle_mstep_nu.restype = C.c_double
        L.oracle_em.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p,
                                _i32p, _f64p, _u8p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.oracle_em.restype = C.c_int
        _LIB = L
    return _LIB


def num_threads() -> int:
    return int(lib().oracle_num_threads())


class OracleProblem:
    """Flattened inputs for the oracle: a FlatForest-like object (``parent``,
    ``left``, ``right``, ``order``, ``n_leaves``), ``leaf_row`` (row of ``codes``
    for every leaf id), integer ``codes`` (rows x K) and the CSR prior."""

    def __init__(self, forest, leaf_row, codes, qptr, qval):
        self.n = int(forest.parent.shape[0])
        self.parent = np.ascontiguousarray(forest.parent, np.int32)
        self.left = np.ascontiguousarray(forest.left, np.int32)
        self.right = np.ascontiguousarray(forest.right, np.int32)
        self.order = np.ascontiguousarray(forest.order, np.int32)
        lr = np.full(self.n, -1, np.int32)
        lr[: len(leaf_row)] = leaf_row
        self.leaf_row = lr
        self.codes = np.ascontiguousarray(codes, np.int32)
        self.K = int(self.codes.shape[1])
        self.n_cells = int(self.codes.shape[0])
        self.qptr = np.ascontiguousarray(qptr, np.int32)
        self.qval = np.ascontiguousarray(qval, np.float64)
        if self.qval.size == 0:
            self.qval = np.zeros(1, np.float64)

    def _tree(self):
        return (self.n, self.K, self.parent, self.left, self.right, self.order, self.leaf_row, self.codes)

    def az_partition(self):
        tag = np.empty((self.n, self.K), np.int32)
        lib().oracle_az_partition(*self._tree(), tag)
        return tag

    def inside(self, d, nu, phi, sites=None, nthreads=1, dump=False):
        s_lo, s_hi = (0, self.K) if sites is None else sites
        d = np.ascontiguousarray(d, np.float64)
        site_llh = np.zeros(self.K, np.float64)
        L0 = L1 = None
        p0 = p1 = None
        if dump:
            L0 = np.zeros((self.n, self.K)); L1 = np.zeros((self.n, self.K))
            p0, p1 = L0.ctypes.data, L1.ctypes.data
        llh = lib().oracle_inside(*self._tree(), self.qptr, self.qval, d, float(nu), float(phi),
                                  s_lo, s_hi, nthreads, site_llh.ctypes.data, p0, p1)
        out = {"llh": float(llh), "site_llh": site_llh}
        if dump:
            out["L0"], out["L1"] = L0, L1
        return out

    def estep(self, d, nu, phi, sites=None, nthreads=1, dump=False):
        s_lo, s_hi = (0, self.K) if sites is None else sites
        d = np.ascontiguousarray(d, np.float64)
        S = np.zeros((5, self.n), np.float64)
        rt = np.zeros(1, np.float64)
        site_llh = np.zeros(self.K, np.float64)
        dump_arr = np.zeros((13, self.n, self.K), np.float64) if dump else None
        llh = lib().oracle_estep(*self._tree(), self.qptr, self.qval, d, float(nu), float(phi),
                                 s_lo, s_hi, nthreads, S, rt, site_llh.ctypes.data,
                                 dump_arr.ctypes.data if dump else None)
        out = {"llh": float(llh), "site_llh": site_llh, "S": S, "Rtilde": float(rt[0])}
        if dump:
            names = ["L0", "L1", "out0", "out1", "A", "X", "post0", "post1", "S0", "S1", "S2", "S3", "S4"]
            out.update({k: dump_arr[i] for i, k in enumerate(names)})
        return out

    def R(self):
        leaf_rows = self.leaf_row[self.left < 0]
        return float((self.codes[leaf_rows] != -1).sum())

    def mstep(self, S, Rtilde, d, nu, free_mask=None, optimize_phi=True, optimize_nu=True, phi=0.0,
              dmin=0.005, dmax=10.0, eps_s=1e-6, eps_nu=1e-5):
        fm = np.ones(self.n, np.uint8) if free_mask is None else np.ascontiguousarray(free_mask, np.uint8)
        S = np.ascontiguousarray(S, np.float64)
        Sn = np.zeros_like(S)
        phi_star = lib().oracle_mstep_stats(self.n, self.K, self.n_cells, S, fm, self.R(), float(Rtilde),
                                            eps_s, Sn)
        if not optimize_phi:
            phi_star = phi
        d = np.array(d, np.float64, copy=True)
        lib().oracle_mstep_brlen(self.n, Sn, fm, float(nu), eps_nu, dmin, dmax, d)
        nu_star = lib().oracle_mstep_nu(self.n, Sn, fm, d) if optimize_nu else nu
        return d, float(nu_star), float(phi_star), Sn

    def em(self, d, nu, phi, free_mask=None, optimize_phi=True, optimize_nu=True, dmin=0.005, dmax=10.0,
           conv_eps=1e-8, max_iter=1000, nthreads=1):
        fm = np.ones(self.n, np.uint8) if free_mask is None else np.ascontiguousarray(free_mask, np.uint8)
        d = np.array(d, np.float64, copy=True)
        nu_a = np.array([nu], np.float64); phi_a = np.array([phi], np.float64); nllh = np.zeros(1)
        iters = lib().oracle_em(self.n, self.K, self.n_cells, self.parent, self.left, self.right, self.order,
                                self.leaf_row, self.codes, self.qptr, self.qval, fm, int(optimize_phi),
                                int(optimize_nu), dmin, dmax, conv_eps, max_iter, nthreads, d, nu_a, phi_a, nllh)
        return {"nllh": float(nllh[0]), "d": d, "nu": float(nu_a[0]), "phi": float(phi_a[0]), "iters": int(iters)}
