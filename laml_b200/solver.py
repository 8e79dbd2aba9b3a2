"""``GPU_EM_solver`` -- the reference's solver-class protocol on the B200 engine.

The reference's extension point for the likelihood / EM path is the solver class
handed to ``Topology_search`` and the runner (``laml_libs/Virtual_solver.py:5-17``;
``Topology_search.py:40-48,299``; ``runner.py:213-223,281-289,391-393``).  This
class implements that protocol -- same constructor, same method names, argument
meaning, return tuples and status strings as ``EM_solver``
(``laml_libs/EM_solver.py:18-527`` on top of ``ML_solver.py:18-336``) -- but every
likelihood, E-step and (unconstrained) M-step runs in the CUDA engine through the
C ABI (``include/pmm_b200.h``).  There is no CPU fallback for the likelihood.

What stays on the host, as in the reference: newick parsing / binarisation,
random initial points (``ML_solver.py:195-213``), fixed-branch marking by LCA
(``ML_solver.py:296-305``), the multi-start bookkeeping of ``optimize``
(``ML_solver.py:256-336``) and, when ``ultra_constr=True``, the equality-
constrained branch-length program (``EM_solver.py:349-366``), which receives the
GPU-computed sufficient statistics.  The reference solves that program with
cvxpy -> MOSEK (commercial; not installable here); this class solves the *same*
program exactly with a tree-structured barrier Newton method on node heights
(``laml_b200/ultrametric.py``).
"""
from __future__ import annotations

import os
from copy import deepcopy
from math import exp, log
from random import random, seed
from typing import Dict, List, Optional

import numpy as np

from .capi import Engine, PMMError
from .flatten import encode_sites, flatten_trees
from .tree import read_tree_newick

# laml_libs/core_constants.py:1-10
min_llh = -float("inf")
eps = 1e-10
conv_eps = 1e-8
dmin = 0.005
dmax = 10
MAX_DEVICE_ULTRA_LEVELS = 1024     # deeper trees take the host solver for the constrained M-step (the device one is level-synchronous)
DEFAULT_STRATEGY = {'resolve_search_only': False, 'only_marked': False, 'ultra_constr': False, 'fixed_phi': None,
                    'fixed_nu': None, 'local_brlen_opt': True, 'nodes_to_recompute': None}


class Params:                                   # ML_solver.py:13-16
    def __init__(self, nu, phi):
        self.nu = nu
        self.phi = phi


def find_LCAs(tree, queries):
    """LCA node of each leaf-label tuple (``laml_libs/lca_lib.py:13-94``; here by
    walking up from the deepest leaf -- queries are few and short)."""
    depth, l2n = {}, {}
    for nd in tree.traverse_preorder():
        depth[nd] = 0 if nd.parent is None else depth[nd.parent] + 1
        if nd.is_leaf():
            l2n[nd.label] = nd
    out = []
    for q in queries:
        nodes = [l2n[a] for a in q if a in l2n]
        if not nodes:
            out.append(None)
            continue
        cur = nodes[0]
        for other in nodes[1:]:
            a, b = cur, other
            while depth[a] > depth[b]:
                a = a.parent
            while depth[b] > depth[a]:
                b = b.parent
            while a is not b:
                a, b = a.parent, b.parent
            cur = a
        out.append(cur)
    return out


class GPU_EM_solver:
    def __init__(self, treeList, data, prior, params={'nu': 0, 'phi': 0, 'sigma': 0}, device: int = 0):
        self.solver_name = "GPU_EM_solver"
        charMtrx = data['charMtrx']
        Q = prior['Q']
        self.trees = []
        for tree in treeList:                                     # ML_solver.py:28-32
            t = read_tree_newick(tree)
            t.suppress_unifurcations()
            self.trees.append(t)
        self.charMtrx = self._reconcile_sites(charMtrx, Q)       # ML_solver.py:47-67
        self.Q = []
        for Q_i in Q:                                             # ML_solver.py:70-74
            s = sum([Q_i[x] for x in Q_i])
            self.Q.append({x: Q_i[x] / s for x in Q_i})
        self.params = Params(params['nu'], params['phi'])
        self.numsites = len(self.charMtrx[next(iter(self.charMtrx.keys()))])
        self.dmin, self.dmax = dmin, dmax
        self.device = device
        self.has_polytomy = False
        self.__mark_polytomies__(eps_len=self.dmin * 0.01)       # EM_solver.py:23
        self.num_edges = sum(len(list(tree.traverse_postorder())) for tree in self.trees)
        for tree in self.trees:
            for node in tree.traverse_postorder():
                node.mark_fixed = False
        self._enc = encode_sites(self.charMtrx, self.Q)
        self._engine: Optional[Engine] = None
        self._leaf_row_bound = None
        self._forest = None
        self._stats = None

    # ------------------------------------------------------------------ set-up
    @staticmethod
    def _reconcile_sites(charMtrx, Q):
        site_matrix = list(zip(*charMtrx.values()))
        if len(Q) != len(site_matrix):
            uninformative = [i for i, site in enumerate(site_matrix) if all(val in (0, -1, '?') for val in site)]
            print(f"Warning: Q has {len(Q)} entries, {len(site_matrix)} sites in charMtrx. "
                  f"Identified {len(uninformative)} uninformative sites.")
            if len(uninformative) == max(len(Q), len(site_matrix)) - min(len(Q), len(site_matrix)):
                keep = [i for i in range(len(site_matrix)) if i not in set(uninformative)]
                charMtrx = {cell: [values[i] for i in keep] for cell, values in charMtrx.items()}
        return charMtrx

    def __mark_polytomies__(self, eps_len=0):                     # EM_solver.py:26-43
        self.has_polytomy = False
        self.num_polytomy_mark = 0
        for tree in self.trees:
            for node in tree.traverse_preorder():
                node.polytomy_mark = False
                if len(node.children) > 2:
                    self.has_polytomy = True
            tree.resolve_polytomies()
        for tree in self.trees:
            for node in tree.traverse_preorder():
                if not hasattr(node, 'polytomy_mark'):
                    node.polytomy_mark = True
                    self.num_polytomy_mark += 1
                    node.edge_length = eps_len
                    self.has_polytomy = True

    def _resolve_new_polytomies(self):
        """``optimize`` ends by re-reading its best trees with the solver-inserted branches collapsed
        (``ML_solver.py:314-333``), so ``self.trees`` may be multifurcating again.  The reference never notices:
        it builds a fresh solver -- whose constructor binarises and marks (``EM_solver.py:19-43``) -- for every
        further call.  This class lives on (one handle per search), so the same binarisation is redone here
        whenever the trees are about to be flattened: nodes that exist keep their marks, the inserted ones get
        ``polytomy_mark`` and the constructor's length."""
        if not any(len(nd.children) > 2 for tree in self.trees for nd in tree.traverse_preorder()):
            return
        for tree in self.trees:
            known = {id(nd) for nd in tree.traverse_preorder()}
            tree.resolve_polytomies()
            for nd in tree.traverse_preorder():
                if id(nd) not in known:
                    nd.polytomy_mark, nd.mark_fixed = True, False
                    nd.edge_length = self.dmin * 0.01
                else:
                    nd.polytomy_mark = bool(getattr(nd, 'polytomy_mark', False))
                    nd.mark_fixed = bool(getattr(nd, 'mark_fixed', False))
        self.has_polytomy = True
        self.num_polytomy_mark = sum(1 for tree in self.trees for nd in tree.traverse_preorder() if nd.polytomy_mark)
        self.num_edges = sum(len(list(tree.traverse_postorder())) for tree in self.trees)

    def _bind(self):
        """(re)build the flat forest and hand it to the engine"""
        self._resolve_new_polytomies()
        row = {c: i for i, c in enumerate(self._enc.cells)}
        try:
            forest = flatten_trees(self.trees, leaf_key=row)      # leaf ids follow the matrix rows: stable across topologies
            leaf_row = np.array([row[l] for l in forest.leaf_labels], np.int64)
        except KeyError as e:
            raise KeyError("leaf %s of the tree has no row in the character matrix" % e)
        if self._engine is None or self._leaf_row_bound is None or not np.array_equal(self._leaf_row_bound, leaf_row):
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(forest.n_leaves, self._enc.qptr, self._enc.qval, self._enc.codes[leaf_row],
                                  device=self.device)
            self._leaf_row_bound = leaf_row
        self._engine.set_tree(forest.parent, forest.left, forest.right)
        self._forest = forest
        self._um_for = None                                       # a cached ultrametric problem belongs to the old topology
        self._R = float((self._enc.codes[leaf_row] != -1).sum())  # EM_solver.py:322
        return forest

    def _ensure(self):
        if self._forest is None:
            self._bind()
        return self._forest

    def _d(self):
        return np.array([float(nd.edge_length) for nd in self._forest.nodes], np.float64)

    def _free_mask(self, local_brlen_opt=True):
        return np.array([0 if (nd.polytomy_mark or (nd.mark_fixed and local_brlen_opt)) else 1
                         for nd in self._forest.nodes], np.uint8)

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    # ------------------------------------------------------------------ protocol
    def get_tree_newick(self):                                    # ML_solver.py:88-89
        return [tree.newick() for tree in self.trees]

    def get_params(self):                                         # ML_solver.py:91-92
        return {'phi': self.params.phi, 'nu': self.params.nu}

    def az_partition(self):
        """The full-state engine needs no z/alpha/? partition (it is implicit in the
        zeros of the state vectors); kept for call compatibility (ML_solver.py:125)."""
        self._ensure()

    def lineage_llh(self):                                        # EM_solver.py:135-141
        self._ensure()
        return self._engine.inside(self._d(), self.params.nu, self.params.phi)

    def negative_llh(self):                                       # ML_solver.py:252-254
        self.az_partition()
        return -self.lineage_llh()

    def Estep(self, posteriors: bool = False):
        """EM_solver.Estep (``EM_solver.py:302-305``).  Leaves ``S0..S4`` (site sums, what
        ``Mstep`` consumes at ``:341``) on every node; with ``posteriors=True`` also the
        per-site ``post0`` / ``post1`` lists the annotation writer reads (``runner.py:418-435``)."""
        self._ensure()
        d = self._d()
        out = self._engine.estep(d, self.params.nu, self.params.phi)
        self._stats = out
        for i, nd in enumerate(self._forest.nodes):
            nd.S0, nd.S1, nd.S2, nd.S3, nd.S4 = (float(x) for x in out["S"][i])
        if posteriors:
            p0, p1 = self._engine.posteriors(d, self.params.nu, self.params.phi)
            for i, nd in enumerate(self._forest.nodes):
                nd.post0, nd.post1 = p0[i].tolist(), p1[i].tolist()
        return out

    # ------------------------------------------------------------------ M-step
    def ultrametric_constr(self, local_brlen_opt=True):
        """Rows of ``M d = b`` over the free branches (``EM_solver.py:53-89``), built
        sparsely: each node carries the set of free-branch indices on the path to one
        of its leaves and the summed length of the fixed branches on it."""
        self._ensure()
        # columns follow the engine's node ids (the order of the free-branch mask), not the
        # reference's postorder; the set of equalities is the same
        free_idx, idx = {}, 0
        for node in self._forest.nodes:
            if not node.polytomy_mark and not (node.mark_fixed and local_brlen_opt):
                free_idx[node] = idx
                idx += 1
        rows, seen = [], set()
        path, const = {}, {}
        for tree in self.trees:
            for node in tree.traverse_postorder():
                if node.is_leaf():
                    p, c = frozenset(), (node.edge_length if node.mark_fixed else 0)
                else:
                    c1, c2 = node.children
                    plus, minus = path[c1] - path[c2], path[c2] - path[c1]
                    key, key_c = (plus, minus), (minus, plus)
                    if (plus or minus) and key not in seen and key_c not in seen:
                        seen.add(key)
                        rows.append((plus, minus, const[c2] - const[c1]))
                    p = path[c1]
                    c = const[c1] + (node.edge_length if node.mark_fixed else 0)
                if node in free_idx:
                    p = p | {free_idx[node]}
                path[node], const[node] = p, c
        r0 = self.trees[0].root
        for tree in self.trees[1:]:
            rows.append((path[r0] - path[tree.root], path[tree.root] - path[r0], const[tree.root] - const[r0]))
        M = np.zeros((len(rows), idx))
        b = np.zeros(len(rows))
        for i, (plus, minus, c) in enumerate(rows):
            M[i, list(plus)] = 1.0
            M[i, list(minus)] = -1.0
            b[i] = c
        return M, b

    def Mstep(self, optimize_phi=True, optimize_nu=True, verbose=1, eps_nu=1e-5, eps_s=1e-6, ultra_constr_cache=None,
              local_brlen_opt=True):
        """EM_solver.Mstep (``EM_solver.py:307-459``).  Without the ultrametric equalities
        everything runs on the GPU; with them the branch-length program runs on the host
        from the GPU's statistics and nu / phi are then solved on the GPU."""
        if self._stats is None:
            raise RuntimeError("Mstep needs a preceding Estep")
        fm = self._free_mask(local_brlen_opt)
        if ultra_constr_cache is not None:
            from . import ultrametric as um
            if not isinstance(ultra_constr_cache, um.UltrametricProblem):
                # the reference hands Mstep what ultrametric_constr() returned, the dense (M, b) (EM_solver.py:470,489);
                # the same equalities in the height parametrisation, built once per such object
                key = id(ultra_constr_cache)
                if getattr(self, "_um_for", None) != (key, bool(local_brlen_opt)):
                    self._um_problem = self._ultrametric_problem(local_brlen_opt)
                    self._um_for = (key, bool(local_brlen_opt))
                ultra_constr_cache = self._um_problem
        d_new, nu_new, phi_new, status_d = self._mstep_arrays(self._stats["S"], self._stats["Rtilde"], fm, self._d(),
                                                              self.params.nu, self.params.phi, optimize_phi, optimize_nu,
                                                              eps_nu, eps_s, ultra_constr_cache)
        if status_d == "infeasible":
            return False, "d_infeasible"                          # EM_solver.py:425-426
        self.params.phi, self.params.nu = phi_new, nu_new
        for i, nd in enumerate(self._forest.nodes):
            if fm[i]:
                nd.edge_length = float(d_new[i])
        success = status_d == "optimal"
        return success, ("optimal" if success else "failed_d")

    def _mstep_arrays(self, S, Rt, fm, d, nu, phi, optimize_phi, optimize_nu, eps_nu, eps_s, cache):
        """one M-step on arrays: ``(d_new, nu_new, phi_new, status_d)``; ``cache`` = the ultrametric problem or None"""
        n_cells = len(self.charMtrx)
        if cache is None:
            d_new, nu_new, phi_new = self._engine.mstep(S, Rt, self._R, n_cells, d, nu, phi, fm,
                                                        optimize_nu, optimize_phi, self.dmin, self.dmax, eps_s, eps_nu)
            return d_new, nu_new, phi_new, "optimal"
        from . import ultrametric as um
        d_new, nu_new, phi_new, H, status_d = um.mstep(cache, S, Rt, self._R, n_cells, self.numsites, fm, d, nu, phi,
                                                       H_prev=getattr(self, "_H_prev", None), optimize_phi=optimize_phi,
                                                       optimize_nu=optimize_nu, eps_nu=eps_nu, eps_s=eps_s)
        if status_d != "infeasible":
            self._H_prev = H
        return d_new, nu_new, phi_new, status_d

    def _ultrametric_problem(self, local_brlen_opt=True):
        """edge kinds for the height re-parametrisation: free / fixed length / counts as zero (polytomy mark),
        exactly the three cases of EM_solver.py:60-79"""
        from . import ultrametric as um
        self._ensure()
        f = self._forest
        kind = np.zeros(f.n_nodes, np.int8)
        fixed_len = np.zeros(f.n_nodes)
        for i, nd in enumerate(f.nodes):
            if nd.polytomy_mark:
                kind[i] = um.ZERO
            elif nd.mark_fixed and local_brlen_opt:
                kind[i] = um.FIXED
                fixed_len[i] = float(nd.edge_length)
        self._H_prev = None
        return um.UltrametricProblem(f.parent, kind, fixed_len, self.dmin, self.dmax)

    # ------------------------------------------------------------------ EM
    def EM_optimization(self, verbose=1, optimize_phi=True, optimize_nu=True, ultra_constr=False, maxIter=1000):
        """EM_solver.EM_optimization (``EM_solver.py:461-512``): returns ``(nllh, em_iter, status)``."""
        self._ensure()
        if not ultra_constr and verbose > 0 and hasattr(self._engine, "em_step"):
            # progress output as the reference's (EM_solver.py:468-505) needs the likelihood of every iterate on the host:
            # one pmm_em_step per iteration -- parameters up, E-step + M-step on the device, parameters and the
            # log-likelihood of the INPUT point down.  Call j turns p_{j-1} into p_j and reports llh(p_{j-1}), so the
            # test of iteration j - 1 (:503) is applied one call later; the converged point is that call's input.
            eng, fm, n_cells = self._engine, self._free_mask(True), len(self.charMtrx)
            cur = (self._d(), self.params.nu, self.params.phi)
            pre_llh, em_iter = None, 0
            while True:
                d_new, nu_new, phi_new, llh = eng.em_step(cur[0], cur[1], cur[2], self._R, n_cells, fm, optimize_nu, optimize_phi,
                                                         self.dmin, self.dmax)
                if pre_llh is None:
                    print("Initial phi: " + str(cur[2]) + ". Initial nu: " + str(cur[1]) + ". Initial nllh: " + str(-llh))
                else:
                    print("Finished EM iter: " + str(em_iter) + ". Current nllh: " + str(-llh))
                    if abs((llh - pre_llh) / pre_llh) < conv_eps or em_iter >= maxIter:
                        if em_iter >= maxIter and not abs((llh - pre_llh) / pre_llh) < conv_eps:
                            em_iter += 1
                            print("Warning: exceeded maximum number of EM iterations (" + str(maxIter) + " iters)!")
                        break
                pre_llh = llh
                em_iter += 1
                cur = (d_new, nu_new, phi_new)
            for nd, length in zip(self._forest.nodes, cur[0].tolist()):
                nd.edge_length = length
            self.params.nu, self.params.phi = cur[1], cur[2]
            return -llh, em_iter, "optimal"
        if not ultra_constr:
            res = self._engine.em(self._d(), self.params.nu, self.params.phi, self._R, len(self.charMtrx),
                                  self._free_mask(True), optimize_nu, optimize_phi, self.dmin, self.dmax, conv_eps, maxIter)
            for i, nd in enumerate(self._forest.nodes):
                nd.edge_length = float(res["d"][i])
            self.params.nu, self.params.phi = res["nu"], res["phi"]
            if verbose >= 0 and res["iters"] > maxIter:
                print("Warning: exceeded maximum number of EM iterations (" + str(maxIter) + " iters)!")
            return res["nllh"], res["iters"], res["status"]
        # ---- ultra_constr: the same loop with the branch-length program under the ultrametric equalities (:349-366)
        eng = self._engine
        d, fm = self._d(), self._free_mask(True)
        nu, phi = self.params.nu, self.params.phi
        cache = self._ultrametric_problem(local_brlen_opt=True)
        n_cells = len(self.charMtrx)
        on_device = (hasattr(eng, "em_ultra") and os.environ.get("PMM_ULTRA_HOST", "0") != "1"
                     and eng.info()["inside_levels"] <= MAX_DEVICE_ULTRA_LEVELS)
        if on_device and verbose <= 0:
            # everything on the device: E-step, the constrained M-step as one cooperative launch (pmm_em_ultra)
            res = eng.em_ultra(d, nu, phi, self._R, n_cells, cache.kind, cache.fixed_len, optimize_nu, optimize_phi,
                               self.dmin, self.dmax, conv_eps, maxIter)
            for nd, length in zip(self._forest.nodes, res["d"].tolist()):
                nd.edge_length = length
            self.params.nu, self.params.phi = res["nu"], res["phi"]
            if res["status"] == "d_infeasible":
                if verbose >= 0:
                    print("Warning: EM failed to optimize parameters in one Mstep due to infeasible constraints")
            elif verbose >= 0 and res["iters"] > maxIter:
                print("Warning: exceeded maximum number of EM iterations (" + str(maxIter) + " iters)!")
            return res["nllh"], res["iters"], res["status"]
        # host-driven loop: progress output (verbose), the CPU stand-in of the tests, very deep trees (the device solver is
        # level-synchronous).  One E-step per iteration: the E-step at p_k also yields llh(p_k), the convergence test of
        # iteration k (the reference runs a separate inside pass for it, :500).  With the engine's device M-step the
        # statistics and parameters stay on the device; otherwise they go through the host solver (laml_b200/ultrametric.py).
        if on_device:
            eng.upload_params(d, nu, phi)
            eng.estep_dev()
            pre_llh = eng.read_llh()
        else:
            est = eng.estep(d, nu, phi)
            self._stats = est
            pre_llh = est["llh"]
        if verbose >= 0:
            print("Initial phi: " + str(phi) + ". Initial nu: " + str(nu) + ". Initial nllh: " + str(-pre_llh))
        em_iter, converged, status, curr_llh = 1, False, "optimal", pre_llh
        infeasible = False
        while em_iter <= maxIter:
            if on_device:
                status_d, _ = eng.mstep_ultra_dev(self._R, n_cells, cache.kind, cache.fixed_len, optimize_nu, optimize_phi,
                                                  self.dmin, self.dmax, warm=em_iter > 1)
            else:
                d_new, nu, phi, status_d = self._mstep_arrays(est["S"], est["Rtilde"], fm, d, nu, phi, optimize_phi, optimize_nu,
                                                              1e-5, 1e-6, cache)
            if status_d == "infeasible":
                infeasible = True
                break
            status = "optimal" if status_d == "optimal" else "failed_d"
            if on_device:
                eng.estep_dev()
                curr_llh = eng.read_llh()
            else:
                d = np.where(fm.astype(bool), d_new, d)
                est = eng.estep(d, nu, phi)
                self._stats = est
                curr_llh = est["llh"]
            if verbose > 0:
                print("Finished EM iter: " + str(em_iter) + ". Current nllh: " + str(-curr_llh))
            if abs((curr_llh - pre_llh) / pre_llh) < conv_eps:
                converged = True
                break
            pre_llh = curr_llh
            em_iter += 1
        if on_device:
            d, nu, phi = eng.download_params()
        self.params.nu, self.params.phi = nu, phi
        for nd, length in zip(self._forest.nodes, d.tolist()):
            nd.edge_length = length
        if infeasible:
            if verbose >= 0:
                print("Warning: EM failed to optimize parameters in one Mstep due to infeasible constraints")
            return -pre_llh, em_iter, "d_infeasible"
        if not converged and verbose >= 0:
            print("Warning: exceeded maximum number of EM iterations (" + str(maxIter) + " iters)!")
        return -curr_llh, em_iter, status

    # -- initial points, ML_solver.py:195-250
    def ini_brlens(self):
        x = [random() * (self.dmax / 2 - 2 * self.dmin) + 2 * self.dmin for _ in range(self.num_edges)]
        idx = 0
        for tree in self.trees:
            for node in tree.traverse_postorder():
                if node.edge_length is not None:
                    x[idx] = max(2 * self.dmin, node.edge_length)
                idx += 1
        return x

    def ini_nu(self, fixed_nu=None):
        return random() * 0.99 if fixed_nu is None else fixed_nu

    def ini_phi(self, fixed_phi=None):
        return random() * 0.99 if fixed_phi is None else fixed_phi

    def ini_all(self, fixed_phi=None, fixed_nu=None):
        return self.ini_brlens() + [self.ini_nu(fixed_nu=fixed_nu), self.ini_phi(fixed_phi=fixed_phi)]

    def x2brlen(self, x):                                         # EM_solver.py:45-51
        i = 0
        for tree in self.trees:
            for node in tree.traverse_postorder():
                if not node.polytomy_mark:
                    node.edge_length = x[i]
                    i += 1

    def x2params(self, x, fixed_nu=None, fixed_phi=None):         # ML_solver.py:240-243
        self.x2brlen(x)
        self.params.nu = x[self.num_edges] if fixed_nu is None else fixed_nu
        self.params.phi = x[self.num_edges + 1] if fixed_phi is None else fixed_phi

    def optimize_one(self, randseed, fixed_phi=None, fixed_nu=None, verbose=1, ultra_constr=False):
        """EM_solver.optimize_one (``EM_solver.py:514-527``)."""
        seed(a=randseed)
        x0 = self.ini_all(fixed_phi=fixed_phi, fixed_nu=fixed_nu)
        self.x2params(x0, fixed_phi=fixed_phi, fixed_nu=fixed_nu)
        self.az_partition()
        nllh, em_iter, status = self.EM_optimization(verbose=verbose, optimize_phi=(fixed_phi is None),
                                                     optimize_nu=(fixed_nu is None), ultra_constr=ultra_constr)
        if verbose >= 0:
            print("EM finished after " + str(em_iter) + " iterations.")
            print("Optimal phi: " + str(self.params.phi) + ". Optimal nu: " + str(self.params.nu) + ". Optimal nllh: " + str(nllh))
        return nllh, status

    def optimize(self, initials=20, fixed_phi=None, fixed_nu=None, fixed_brlen=None, verbose=1, max_trials=100,
                 random_seeds=None, ultra_constr=False):
        """ML_solver.optimize (``ML_solver.py:256-336``): multi-start EM; returns ``(nllh, status)`` or None."""
        results = []
        all_failed, all_trials = True, 0
        if random_seeds is None:
            rseeds = [int(random() * 10000) for _ in range(initials)]
        elif type(random_seeds) == int:
            if verbose >= 0:
                print("Global random seed: " + str(random_seeds))
            seed(a=random_seeds)
            rseeds = [int(random() * 10000) for _ in range(initials)]
        elif type(random_seeds) == list:
            if len(random_seeds) < initials:
                if verbose >= 0:
                    print("Fatal: the number of random seeds is smaller than the number of initials!")
                return None
            rseeds = random_seeds[:initials]
        else:
            if verbose >= 0:
                print("Fatal: incorrect random_seeds type provided")
            return None
        self._ensure()
        while all_failed and all_trials < max_trials:
            for rep in range(initials):
                randseed = rseeds[rep] + all_trials
                if verbose >= 0:
                    print("Initial point " + str(rep + 1) + ". Random seed: " + str(randseed))
                for t, tree in enumerate(self.trees):               # ML_solver.py:296-305
                    for node in tree.traverse_postorder():
                        node.mark_fixed = False
                    if fixed_brlen is None:
                        continue
                    fixed_nodes = find_LCAs(tree, list(fixed_brlen[t].keys()))
                    for i, (a, b) in enumerate(fixed_brlen[t]):
                        u = fixed_nodes[i]
                        u.edge_length = fixed_brlen[t][(a, b)]
                        u.mark_fixed = True
                nllh, status = self.optimize_one(randseed, fixed_phi=fixed_phi, fixed_nu=fixed_nu, verbose=verbose,
                                                 ultra_constr=ultra_constr)
                if nllh is not None:
                    all_failed = False
                    processed = []
                    for tree in self.trees:                         # ML_solver.py:314-319
                        tree_copy = read_tree_newick(tree.newick())
                        tree_copy.collapse_short_branches(self.dmin * 0.01)
                        processed.append(tree_copy.newick())
                    results.append((nllh, rep, deepcopy(self.params), processed, status))
                elif verbose >= 0:
                    print("Fatal: failed to optimize using initial point " + str(rep + 1))
            all_trials += initials
        if all_failed:
            if verbose >= 0:
                print("Fatal: Optimization failed on more than " + str(max_trials) + " retries. Aborting ...")
            return None
        results.sort(key=lambda r: (r[0], r[1]))
        best_nllh, _, best_params, best_trees, status = results[0]
        self.trees = [read_tree_newick(t) for t in best_trees]      # ML_solver.py:329-333
        for tree in self.trees:
            for node in tree.traverse_postorder():
                node.mark_fixed = False
                node.polytomy_mark = False
        self._forest = None
        self.params = best_params
        if verbose >= 0:
            print("Numerical optimization finished successfully")
        return best_nllh, status

    def score_tree(self, strategy={'ultra_constr': False, 'fixed_phi': None, 'fixed_nu': None, 'fixed_brlen': None}):
        """ML_solver.score_tree (``ML_solver.py:114-123``): ``(score = -nllh | None, status)``."""
        out = self.optimize(initials=1, verbose=-1, ultra_constr=strategy['ultra_constr'], fixed_phi=strategy['fixed_phi'],
                            fixed_nu=strategy['fixed_nu'], fixed_brlen=strategy.get('fixed_brlen'))
        if out is None:
            return None, "failed"
        nllh, status = out
        return (None if nllh is None else -nllh), status
