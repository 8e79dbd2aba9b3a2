"""ctypes binding of the C ABI in ``include/pmm_b200.h`` (``libpmm_b200.so``).

This is deliberately thin: numpy arrays in, numpy arrays out, every error from
the library raised as :class:`PMMError`.  There is no fallback of any kind -- if
the shared library is missing or no CUDA device is usable the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMM_LIB") or os.path.join(_HERE, "libpmm_b200.so")   # PMM_LIB: tuning builds only

#: every symbol ``include/pmm_b200.h`` declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "pmm_abi_version", "pmm_last_error", "pmm_create", "pmm_destroy", "pmm_set_stream", "pmm_set_tree",
    "pmm_inside", "pmm_estep", "pmm_posteriors", "pmm_mstep", "pmm_em", "pmm_set_total_sites",
    "pmm_upload_params", "pmm_download_params", "pmm_estep_dev", "pmm_inside_dev", "pmm_stats_dev_ptr",
    "pmm_stats_len", "pmm_mstep_dev", "pmm_sync", "pmm_get_info", "pmm_last_timing", "pmm_score_nni_batch", "pmm_set_option", "pmm_debug_lumped_schedule",
    "pmm_ultrametric_solve", "pmm_score_nni_batch_ultra", "pmm_debug_nni_ultra_mstep", "pmm_em_step",
    "pmm_mstep_ultra_dev", "pmm_em_ultra", "pmm_read_llh", "pmm_debug_lumped_deps",
    "pmm_peer_export", "pmm_peer_setup", "pmm_peer_setup_local", "pmm_peer_close",
]
PEER_HANDLE_BYTES = 64


def nni_ultra_mstep(S25, nu, lo, hi, hA, hB, hC, hTop, top_free=False, H=None, eps_nu=1e-5):
    """``pmm_debug_nni_ultra_mstep`` (host instance of the batch kernel's constrained M-step; no GPU needed):
    ``(d5, H3, status, newton_iters)``; ``H`` = the previous heights (warm start) or None."""
    L = load_library()
    S = np.ascontiguousarray(S25, np.float64).reshape(25)
    H3 = np.zeros(3) if H is None else np.array(H, np.float64, copy=True)
    d5 = np.zeros(5)
    it = C.c_int32()
    rc = L.pmm_debug_nni_ultra_mstep(S.ctypes.data, float(nu), float(eps_nu), float(lo), float(hi), float(hA), float(hB),
                                     float(hC), float(hTop), int(bool(top_free)), int(H is not None), H3.ctypes.data,
                                     d5.ctypes.data, C.byref(it))
    return d5, H3, rc, it.value


class PMMError(RuntimeError):
    pass


class PMMInfo(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "n_nodes", "n_leaves", "n_sites", "slots_per_node", "row_stride", "padded_slots_per_node", "n_chunks", "n_tiles", "tile_slots",
        "inside_levels", "outside_levels", "device_bytes", "kernel_launches", "engine", "clades", "clade_target",
        "launches_per_pass", "inside_bytes_per_site", "outside_bytes_per_site")]


_lib = None


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PMMError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). The PMM engine has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, f64, i64 = C.c_void_p, C.c_int32, C.c_double, C.c_int64
    P = C.POINTER
    L.pmm_abi_version.restype = C.c_int
    L.pmm_last_error.restype = C.c_char_p
    L.pmm_last_error.argtypes = [vp]
    L.pmm_create.argtypes = [i32, i32, vp, vp, vp, i32, C.c_size_t, P(vp)]
    L.pmm_destroy.argtypes = [vp]
    L.pmm_destroy.restype = None
    L.pmm_set_stream.argtypes = [vp, vp]
    L.pmm_set_tree.argtypes = [vp, i32, vp, vp, vp]
    L.pmm_inside.argtypes = [vp, vp, f64, f64, vp, P(f64)]
    L.pmm_estep.argtypes = [vp, vp, f64, f64, vp, P(f64), P(f64), vp]
    L.pmm_posteriors.argtypes = [vp, vp, f64, f64, vp, vp]
    L.pmm_mstep.argtypes = [vp, vp, f64, f64, i32, vp, i32, i32, f64, f64, f64, f64, vp, P(f64), P(f64)]
    L.pmm_em.argtypes = [vp, vp, P(f64), P(f64), vp, i32, i32, f64, i32, f64, f64, f64, i32, P(f64), P(i32), P(i32)]
    L.pmm_em_step.argtypes = [vp, vp, P(f64), P(f64), vp, i32, i32, f64, i32, f64, f64, P(f64)]
    L.pmm_mstep_ultra_dev.argtypes = [vp, f64, i32, vp, vp, i32, i32, f64, f64, f64, f64, i32, P(i32), P(i32)]
    L.pmm_em_ultra.argtypes = [vp, vp, P(f64), P(f64), vp, vp, i32, i32, f64, i32, f64, f64, f64, i32, P(f64), P(i32), P(i32)]
    L.pmm_set_total_sites.argtypes = [vp, i32]
    L.pmm_upload_params.argtypes = [vp, vp, f64, f64]
    L.pmm_download_params.argtypes = [vp, vp, P(f64), P(f64)]
    L.pmm_estep_dev.argtypes = [vp]
    L.pmm_inside_dev.argtypes = [vp]
    L.pmm_stats_dev_ptr.argtypes = [vp]
    L.pmm_stats_dev_ptr.restype = vp
    L.pmm_stats_len.argtypes = [vp]
    L.pmm_stats_len.restype = i64
    L.pmm_mstep_dev.argtypes = [vp, f64, i32, vp, i32, i32, f64, f64, f64, f64]
    L.pmm_score_nni_batch.argtypes = [vp, i32, vp, vp, vp, f64, f64, f64, f64, f64, i32, vp, vp, vp, vp]
    L.pmm_score_nni_batch_ultra.argtypes = [vp, i32, vp, vp, vp, f64, f64, f64, f64, f64, i32, vp, vp, vp, vp]
    L.pmm_debug_nni_ultra_mstep.argtypes = [vp, f64, f64, f64, f64, f64, f64, f64, f64, i32, i32, vp, vp, P(i32)]
    L.pmm_sync.argtypes = [vp]
    L.pmm_read_llh.argtypes = [vp, P(f64)]
    L.pmm_set_option.argtypes = [vp, C.c_char_p, i64]
    L.pmm_debug_lumped_schedule.argtypes = [i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, P(i32), P(i32)]
    L.pmm_debug_lumped_deps.argtypes = [i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.pmm_peer_export.argtypes = [vp, vp]
    L.pmm_peer_setup.argtypes = [vp, i32, i32, vp]
    L.pmm_peer_setup_local.argtypes = [vp, i32, i32, vp]
    L.pmm_peer_close.argtypes = [vp]
    L.pmm_ultrametric_solve.argtypes = [i32, vp, vp, vp, f64, f64, vp, f64, f64, vp, f64, f64, f64, vp, vp, P(i32)]
    L.pmm_get_info.argtypes = [vp, P(PMMInfo)]
    L.pmm_last_timing.argtypes = [vp, P(f64), P(f64), P(f64)]
    _lib = L
    return L


def ultrametric_solve(parent, kind, fixed_len, lo, hi, S, nu, eps_nu=1e-5, H0=None, mu_final=1e-11, warm_theta=1e-6,
                      warm_mu=1e-9):
    """``pmm_ultrametric_solve`` (host code of the library; no GPU needed): ``(d, H, status, newton_iters)`` with
    status 0 optimal / 1 infeasible / 2 failed."""
    L = load_library()
    parent = _i32(parent)
    n = int(parent.shape[0])
    kind = np.ascontiguousarray(kind, np.int8)
    fixed_len = _f64(fixed_len)
    S = _f64(S)
    if kind.shape[0] != n or fixed_len.shape[0] != n or S.shape != (n, 5):
        raise PMMError("pmm_ultrametric_solve: kind, fixed_len must have n entries and S must be (n, 5)")
    h0 = _f64(H0) if H0 is not None else None
    if h0 is not None and h0.shape[0] != n + 1:
        raise PMMError("pmm_ultrametric_solve: H0 must have n + 1 entries")
    d = np.empty(n); H = np.empty(n + 1); it = C.c_int32()
    rc = L.pmm_ultrametric_solve(n, parent.ctypes.data, kind.ctypes.data, fixed_len.ctypes.data, float(lo), float(hi),
                                 S.ctypes.data, float(nu), float(eps_nu), h0.ctypes.data if h0 is not None else None,
                                 float(mu_final), float(warm_theta), float(warm_mu), d.ctypes.data, H.ctypes.data, C.byref(it))
    if rc < 0:
        raise PMMError("pmm_ultrametric_solve: invalid arguments (parent must describe a forest)")
    return d, H, int(rc), int(it.value)


def lumped_schedule(n_leaves, parent, left, right, clade_cap=16, clade_cap_top=16):
    """The clade schedule of the lumped-state engine for a forest (``pmm_debug_lumped_schedule``; pure host code,
    no GPU needed).  Returns a dict of int32 arrays: ``in_steps`` / ``out_steps`` ``(n_int, 4)`` = (u, v, w, flags),
    ``in_clades`` / ``out_clades`` ``(n_clades, 4)`` = (first step, end step, tree index, clade id) and the
    ``*_phase_ptr`` arrays (first clade of every launch)."""
    L = load_library()
    parent, left, right = _i32(parent), _i32(left), _i32(right)
    n = int(parent.shape[0]); n_int = n - int(n_leaves)
    ins, outs = np.zeros((n_int, 4), np.int32), np.zeros((n_int, 4), np.int32)
    inc, outc = np.zeros((n_int, 4), np.int32), np.zeros((n_int, 4), np.int32)
    ipp, opp = np.zeros(n_int + 2, np.int32), np.zeros(n_int + 2, np.int32)
    ncl, nph = C.c_int32(), C.c_int32()
    rc = L.pmm_debug_lumped_schedule(int(n_leaves), n, parent.ctypes.data, left.ctypes.data, right.ctypes.data, int(clade_cap),
                                     int(clade_cap_top), ins.ctypes.data, outs.ctypes.data, inc.ctypes.data, outc.ctypes.data,
                                     ipp.ctypes.data, opp.ctypes.data, C.byref(ncl), C.byref(nph))
    if rc:
        raise PMMError("pmm_debug_lumped_schedule: invalid tree arrays")
    k, p = ncl.value, nph.value
    # the waits of the dataflow launches: per clade position, the positions of its producer clades (CSR)
    idp, odp = np.zeros(n_int + 1, np.int32), np.zeros(n_int + 1, np.int32)
    idd, odd = np.zeros(2 * n_int + 1, np.int32), np.zeros(n_int + 1, np.int32)
    rc = L.pmm_debug_lumped_deps(int(n_leaves), n, parent.ctypes.data, left.ctypes.data, right.ctypes.data, int(clade_cap),
                                 int(clade_cap_top), idp.ctypes.data, idd.ctypes.data, odp.ctypes.data, odd.ctypes.data)
    if rc:
        raise PMMError("pmm_debug_lumped_deps: invalid tree arrays")
    return {"in_steps": ins, "out_steps": outs, "in_clades": inc[:k], "out_clades": outc[:k],
            "in_phase_ptr": ipp[:p + 1], "out_phase_ptr": opp[:p + 1], "n_clades": k, "n_phases": p,
            "in_dep_ptr": idp[:k + 1], "in_deps": idd[:idp[k]], "out_dep_ptr": odp[:k + 1], "out_deps": odd[:odp[k]]}


class Engine:
    """One handle = one GPU, one data set (or site shard), one current tree."""

    def __init__(self, n_leaves: int, qptr, qval, codes, device: int = 0, mem_budget_bytes: int = 0):
        L = load_library()
        self._L = L
        self._h = C.c_void_p()
        qptr = _i32(qptr)
        qval = _f64(qval) if len(qval) else np.zeros(1)
        codes = _i32(codes)
        if codes.ndim != 2 or codes.shape[0] != n_leaves:
            raise PMMError("codes must be (n_leaves, n_sites) in leaf-id order")
        self.n_leaves, self.n_sites = int(n_leaves), int(codes.shape[1])
        rc = L.pmm_create(self.n_leaves, self.n_sites, qptr.ctypes.data, qval.ctypes.data, codes.ctypes.data,
                          int(device), int(mem_budget_bytes), C.byref(self._h))
        if rc:
            raise PMMError(L.pmm_last_error(None).decode())
        self.n_nodes = 0
        self.device = int(device)

    # -- plumbing -----------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise PMMError(self._L.pmm_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.pmm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self._L.pmm_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_tree(self, parent, left, right):
        parent, left, right = _i32(parent), _i32(left), _i32(right)
        self._ck(self._L.pmm_set_tree(self._h, parent.shape[0], parent.ctypes.data, left.ctypes.data, right.ctypes.data))
        self.n_nodes = int(parent.shape[0])

    def set_option(self, name: str, value: int):
        """runtime options of the handle (``include/pmm_b200.h``: "slices")"""
        self._ck(self._L.pmm_set_option(self._h, name.encode(), int(value)))

    def set_total_sites(self, k: int):
        self._ck(self._L.pmm_set_total_sites(self._h, int(k)))

    # -- host-buffer entry points --------------------------------------------
    def inside(self, d, nu, phi, want_site_llh=False):
        d = _f64(d)
        llh = C.c_double()
        site = np.empty(self.n_sites) if want_site_llh else None
        self._ck(self._L.pmm_inside(self._h, d.ctypes.data, float(nu), float(phi),
                                    site.ctypes.data if want_site_llh else None, C.byref(llh)))
        return (llh.value, site) if want_site_llh else llh.value

    def estep(self, d, nu, phi, want_site_llh=False, out_S=None):
        """``out_S``: optional caller-owned (n_nodes, 5) float64 buffer (e.g. pinned)."""
        d = _f64(d)
        S = np.empty((self.n_nodes, 5)) if out_S is None else out_S
        rt, llh = C.c_double(), C.c_double()
        site = np.empty(self.n_sites) if want_site_llh else None
        self._ck(self._L.pmm_estep(self._h, d.ctypes.data, float(nu), float(phi), S.ctypes.data, C.byref(rt),
                                   C.byref(llh), site.ctypes.data if want_site_llh else None))
        out = {"S": S, "Rtilde": rt.value, "llh": llh.value}
        if want_site_llh:
            out["site_llh"] = site
        return out

    def posteriors(self, d, nu, phi):
        d = _f64(d)
        p0 = np.empty((self.n_nodes, self.n_sites)); p1 = np.empty((self.n_nodes, self.n_sites))
        self._ck(self._L.pmm_posteriors(self._h, d.ctypes.data, float(nu), float(phi), p0.ctypes.data, p1.ctypes.data))
        return p0, p1

    def mstep(self, S, Rtilde, R, n_cells, d, nu, phi, free_mask=None, optimize_nu=True, optimize_phi=True,
              dmin=0.005, dmax=10.0, eps_s=1e-6, eps_nu=1e-5):
        S = _f64(S); d = np.array(d, np.float64, copy=True)
        fm = np.ascontiguousarray(free_mask, np.uint8) if free_mask is not None else None
        nu_c, phi_c = C.c_double(nu), C.c_double(phi)
        self._ck(self._L.pmm_mstep(self._h, S.ctypes.data, float(Rtilde), float(R), int(n_cells),
                                   fm.ctypes.data if fm is not None else None, int(optimize_nu), int(optimize_phi),
                                   dmin, dmax, eps_s, eps_nu, d.ctypes.data, C.byref(nu_c), C.byref(phi_c)))
        return d, nu_c.value, phi_c.value

    def em(self, d, nu, phi, R, n_cells, free_mask=None, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0,
           conv_eps=1e-8, max_iter=1000):
        d = np.array(d, np.float64, copy=True)
        fm = np.ascontiguousarray(free_mask, np.uint8) if free_mask is not None else None
        nu_c, phi_c, nllh = C.c_double(nu), C.c_double(phi), C.c_double()
        iters, status = C.c_int32(), C.c_int32()
        self._ck(self._L.pmm_em(self._h, d.ctypes.data, C.byref(nu_c), C.byref(phi_c),
                                fm.ctypes.data if fm is not None else None, int(optimize_nu), int(optimize_phi),
                                float(R), int(n_cells), dmin, dmax, conv_eps, int(max_iter), C.byref(nllh),
                                C.byref(iters), C.byref(status)))
        return {"nllh": nllh.value, "d": d, "nu": nu_c.value, "phi": phi_c.value, "iters": iters.value,
                "status": "optimal" if status.value == 0 else "failed"}

    def em_step(self, d, nu, phi, R, n_cells, free_mask=None, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0, out=None):
        """``pmm_em_step``: one EM iteration, host parameters in / out: ``(d, nu, phi, llh at the input point)``.
        ``out``: a float64 array for the new lengths (may be ``d`` itself); if it is page-locked (``torch``'s
        ``pin_memory``), the copy engines read and write it in place -- no staging copy in either direction."""
        if out is None:
            d = np.array(d, np.float64, copy=True)
        else:
            if out is not d:
                np.copyto(out, d)
            d = out
        fm = np.ascontiguousarray(free_mask, np.uint8) if free_mask is not None else None
        nu_c, phi_c, llh = C.c_double(nu), C.c_double(phi), C.c_double()
        self._ck(self._L.pmm_em_step(self._h, d.ctypes.data, C.byref(nu_c), C.byref(phi_c),
                                     fm.ctypes.data if fm is not None else None, int(optimize_nu), int(optimize_phi),
                                     float(R), int(n_cells), dmin, dmax, C.byref(llh)))
        return d, nu_c.value, phi_c.value, llh.value

    def score_nni_batch(self, cand_u, cand_k, d, nu, phi, dmin=0.005, dmax=10.0, conv_eps=1e-8, max_iter=1000, ultra=False):
        """Score NNI candidates (``Topology_search.apply_nni``, one local 5-branch EM each) in one call.
        ``ultra=True``: the local M-steps under the ultrametric constraint (``pmm_score_nni_batch_ultra``);
        ``d`` must be ultrametric; candidates the constraint leaves no room for come back with ``nllh`` NaN
        and ``iters`` -1."""
        cu, ck = _i32(cand_u), _i32(cand_k)
        m = int(cu.shape[0])
        d = _f64(d)
        nllh = np.empty(m); d5 = np.empty((m, 5)); iters = np.empty(m, np.int32); nodes = np.empty((m, 5), np.int32)
        fn = self._L.pmm_score_nni_batch_ultra if ultra else self._L.pmm_score_nni_batch
        self._ck(fn(self._h, m, cu.ctypes.data, ck.ctypes.data, d.ctypes.data, float(nu), float(phi),
                    dmin, dmax, conv_eps, int(max_iter), nllh.ctypes.data, d5.ctypes.data,
                    iters.ctypes.data, nodes.ctypes.data))
        return {"nllh": nllh, "d5": d5, "iters": iters, "nodes": nodes}

    # -- device-resident entry points ----------------------------------------
    def upload_params(self, d, nu, phi):
        d = _f64(d)
        self._ck(self._L.pmm_upload_params(self._h, d.ctypes.data, float(nu), float(phi)))

    def download_params(self):
        d = np.empty(self.n_nodes)
        nu, phi = C.c_double(), C.c_double()
        self._ck(self._L.pmm_download_params(self._h, d.ctypes.data, C.byref(nu), C.byref(phi)))
        return d, nu.value, phi.value

    def estep_dev(self):
        self._ck(self._L.pmm_estep_dev(self._h))

    def inside_dev(self):
        self._ck(self._L.pmm_inside_dev(self._h))

    def mstep_dev(self, R, n_cells, free_mask=None, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0,
                  eps_s=1e-6, eps_nu=1e-5):
        fm = np.ascontiguousarray(free_mask, np.uint8) if free_mask is not None else None
        self._ck(self._L.pmm_mstep_dev(self._h, float(R), int(n_cells), fm.ctypes.data if fm is not None else None,
                                       int(optimize_nu), int(optimize_phi), dmin, dmax, eps_s, eps_nu))

    UM_STATUS = {0: "optimal", 1: "infeasible", 2: "failed_d"}

    def mstep_ultra_dev(self, R, n_cells, kind, fixed_len, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0,
                        eps_s=1e-6, eps_nu=1e-5, warm=True):
        """``pmm_mstep_ultra_dev``: the ultrametric-constrained M-step on the statistics the device holds;
        returns ``(status, newton_iters)``; the parameters stay on the device (``download_params``)."""
        kind = np.ascontiguousarray(kind, np.int8); fixed_len = _f64(fixed_len)
        st, it = C.c_int32(), C.c_int32()
        self._ck(self._L.pmm_mstep_ultra_dev(self._h, float(R), int(n_cells), kind.ctypes.data, fixed_len.ctypes.data,
                                             int(optimize_nu), int(optimize_phi), dmin, dmax, eps_s, eps_nu, int(bool(warm)),
                                             C.byref(st), C.byref(it)))
        return self.UM_STATUS[st.value], it.value

    def em_ultra(self, d, nu, phi, R, n_cells, kind, fixed_len, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0,
                 conv_eps=1e-8, max_iter=1000):
        """``pmm_em_ultra``: EM under the ultrametric constraint, device-resident"""
        d = np.array(d, np.float64, copy=True)
        kind = np.ascontiguousarray(kind, np.int8); fixed_len = _f64(fixed_len)
        nu_c, phi_c, nllh = C.c_double(nu), C.c_double(phi), C.c_double()
        iters, status = C.c_int32(), C.c_int32()
        self._ck(self._L.pmm_em_ultra(self._h, d.ctypes.data, C.byref(nu_c), C.byref(phi_c), kind.ctypes.data, fixed_len.ctypes.data,
                                      int(optimize_nu), int(optimize_phi), float(R), int(n_cells), dmin, dmax, conv_eps,
                                      int(max_iter), C.byref(nllh), C.byref(iters), C.byref(status)))
        return {"nllh": nllh.value, "d": d, "nu": nu_c.value, "phi": phi_c.value, "iters": iters.value,
                "status": {0: "optimal", 1: "d_infeasible", 2: "failed_d"}[status.value]}

    # ---- the exchange over peer memory (site shards on several GPUs): after ``peer_setup`` every E-step / inside pass
    # leaves the statistics all-reduced over the ranks, computed by the finishing kernel itself over NVLink
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        self._ck(self._L.pmm_peer_export(self._h, buf))
        return buf.raw

    def peer_setup(self, rank: int, world: int, handles: bytes):
        """``handles``: the ranks' ``peer_export()`` results concatenated in rank order"""
        if len(handles) != world * PEER_HANDLE_BYTES:
            raise PMMError("peer_setup: need %d bytes of handles, got %d" % (world * PEER_HANDLE_BYTES, len(handles)))
        self._ck(self._L.pmm_peer_setup(self._h, int(rank), int(world), C.c_char_p(handles)))

    def peer_setup_local(self, rank: int, engines):
        """ranks that are handles of this process (tests; several GPUs driven by one process)"""
        arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
        self._ck(self._L.pmm_peer_setup_local(self._h, int(rank), len(engines), arr))

    def peer_close(self):
        if self._h:
            self._ck(self._L.pmm_peer_close(self._h))

    def stats_ptr(self):
        return int(self._L.pmm_stats_dev_ptr(self._h)), int(self._L.pmm_stats_len(self._h))

    def read_llh(self) -> float:
        """the log-likelihood the last ``estep_dev`` / ``inside_dev`` left on the device"""
        x = C.c_double()
        self._ck(self._L.pmm_read_llh(self._h, C.byref(x)))
        return x.value

    def sync(self):
        self._ck(self._L.pmm_sync(self._h))

    def info(self) -> dict:
        inf = PMMInfo()
        self._ck(self._L.pmm_get_info(self._h, C.byref(inf)))
        return {k: int(getattr(inf, k)) for k, _ in PMMInfo._fields_}

    def last_timing(self):
        a, b, t = C.c_double(), C.c_double(), C.c_double()
        self._ck(self._L.pmm_last_timing(self._h, C.byref(a), C.byref(b), C.byref(t)))
        return {"inside_ms": a.value, "outside_ms": b.value, "total_ms": t.value}
