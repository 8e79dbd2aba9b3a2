"""NNI topology search on the batched scorer: the tree-level side of
``pmm_score_nni_batch``.

The reference's ``Topology_search`` (``laml_libs/Topology_search.py``) tries candidates one at a time:
for a branch ``u`` it re-wires the tree objects, serialises them to newick, constructs a NEW solver from
the strings (re-parsing the trees, re-normalising the prior, re-encoding the character matrix), finds the
LCAs of every fixed branch to pin its length, runs a local EM with five free branches, and accepts or
reverts (``apply_nni``, ``:242-322``).  Here the solver object and its GPU handle live for the whole
search; one call scores EVERY candidate around the current tree (:func:`score_neighbourhood`, 2(N-2)
local EMs in one launch), the acceptance policy -- the reference's own: random branch order, the two
moves of a branch in random order, simulated-annealing acceptance (``:67-80,221-240``) -- is replayed on
the host over the returned scores, and an accepted move is a pointer swap on the tree objects plus a
``pmm_set_tree`` on the live handle (:func:`apply_move`).

What is deliberately NOT reproduced: the reference's random-number stream (every ``score_tree`` call
re-seeds the global generator, ``ML_solver.py:273-276,EM_solver.py:516``, so its trajectory depends on
how many candidates were scored) and the stale branch lengths it keeps on ``treeList_obj`` after an
accepted move (``Topology_search.py:289``: "possibility 1 ..."); here the accepted lengths are the
current lengths.  Scores of individual candidates are identical to the reference's call
``solver(newick of the re-wired tree).score_tree(strategy with fixed_brlen / fixed_nu / fixed_phi)``
(``tests/test_topology_host.py``, ``tests/test_zz_gpu_topology.py``).
"""
from __future__ import annotations

from copy import deepcopy
from dataclasses import dataclass
from math import exp
from random import getrandbits, random, shuffle
from typing import List, Optional

import numpy as np

from .flatten import forest_from_parent
from .nni import apply_nni, nni_candidates
from .solver import DEFAULT_STRATEGY, GPU_EM_solver

chkpt_freq = 10                                   # laml_libs/core_constants.py


@dataclass
class Neighbourhood:
    """Scores of the NNI candidates around a solver's current forest.  Candidate ``i`` swaps child
    ``cand_k[i]`` (index into ``u.children``) of node ``cand_u[i]`` with ``u``'s sibling; ``score[i]`` is the
    log-likelihood after the local EM, ``d5[i]`` the optimised lengths of ``nodes[i]`` = (v, u, moved child,
    other child, former sibling); ``d0`` the lengths every other branch was held at (``d_raw`` raised to
    ``2 dmin``)."""
    cand_u: np.ndarray
    cand_k: np.ndarray
    score: np.ndarray
    d5: np.ndarray
    nodes: np.ndarray
    iters: np.ndarray
    d0: np.ndarray
    d_raw: np.ndarray

    def index(self):
        return {(int(u), int(k)): i for i, (u, k) in enumerate(zip(self.cand_u, self.cand_k))}

    def first_of(self, n_nodes: int) -> np.ndarray:
        """``pos[u]`` = index of candidate ``(u, 0)`` (``(u, 1)`` follows it), -1 for nodes without candidates"""
        pos = np.full(n_nodes, -1, np.int64)
        pos[self.cand_u[::2]] = np.arange(0, len(self.cand_u), 2)
        return pos


def start_lengths(solver: GPU_EM_solver) -> np.ndarray:
    """The point the reference's candidate EM starts from: every length raised to ``2 dmin``
    (``ini_brlens``, ``ML_solver.py:195-203``, applied to fixed and free branches alike by ``x2brlen``)."""
    solver._ensure()
    return np.maximum(solver._d(), 2 * solver.dmin)


def is_ultrametric(forest, d, tol: float = 1e-9) -> bool:
    """leaf paths below every node equal to ``tol`` and all trees equally high: what ``pmm_score_nni_batch_ultra``
    requires of the current lengths (everything outside a candidate's five branches is a constant of its problem)"""
    n = forest.n_nodes
    parent = np.asarray(forest.parent)
    depth = np.zeros(n)                                            # root-to-node path including the root branch
    order = [int(r) for r in np.nonzero(parent < 0)[0]]
    for v in order:                                                # parents before children
        depth[v] = d[v] + (depth[parent[v]] if parent[v] >= 0 else 0.0)
        if forest.left[v] >= 0:
            order += [int(forest.left[v]), int(forest.right[v])]
    tips = depth[: forest.n_leaves]
    return bool(tips.max() - tips.min() <= tol)


def score_neighbourhood(solver: GPU_EM_solver, only_marked: bool = False, clamp: bool = True, shard: bool = False,
                        group=None, device=None, ultra: bool = False) -> Neighbourhood:
    """One batched call for all candidates ``Topology_search.single_nni`` could try on ``solver``'s trees
    (``:221-231``: every internal non-root node; with ``only_marked`` those carrying ``node.mark``), each
    scored as ``apply_nni`` does with ``local_brlen_opt`` (``:253-300``): nu, phi and all lengths fixed
    except the five around the move.

    ``clamp=True`` reproduces the reference's scores exactly: its candidate solver raises EVERY length --
    the fixed ones too -- to ``2 dmin`` before the local EM (:func:`start_lengths`).  With ``clamp=False`` the
    other branches stay where they are (see :class:`GPU_Topology_search`).

    ``shard=True`` under ``torch.distributed``: candidate trees are independent units, so every rank -- each
    holding all sites -- scores the candidates ``rank::world`` and one ``all_gather`` gives every rank the full
    result (``dist.score_nni_sharded``, BASELINE config 5 on 8 GPUs).

    ``ultra=True``: the local EM of every candidate under the ultrametric constraint
    (``pmm_score_nni_batch_ultra``); candidates it leaves no room for have score NaN."""
    forest = solver._ensure()
    cu, ck = nni_candidates(forest.parent, forest.left, forest.n_leaves)
    if only_marked:
        keep = np.array([bool(getattr(forest.nodes[u], "mark", False)) for u in cu], bool)
        cu, ck = cu[keep], ck[keep]
    if solver.num_polytomy_mark:
        # ``search(resolve_polytomies=False)`` on a polytomous input (runner.py:312, --keep_polytomies): the solver
        # binarised the trees itself and its inserted branches are never optimised (EM_solver.py:39-43), so a move
        # whose five branches include one is not a candidate -- the kernel would treat it as a free branch
        pm = np.array([bool(getattr(nd, "polytomy_mark", False)) for nd in forest.nodes], bool)
        if pm.any():
            v = forest.parent[cu]
            w = np.where(forest.left[v] == cu, forest.right[v], forest.left[v])
            keep = ~(pm[cu] | pm[v] | pm[w] | pm[forest.left[cu]] | pm[forest.right[cu]])
            cu, ck = cu[keep], ck[keep]
    d_raw = solver._d()
    d0 = np.maximum(d_raw, 2 * solver.dmin) if clamp else d_raw.copy()
    if len(cu) == 0:
        z = np.zeros(0)
        return Neighbourhood(cu, ck, z, np.zeros((0, 5)), np.zeros((0, 5), np.int32), np.zeros(0, np.int32), d0, d_raw)
    kw = {"ultra": True} if ultra else {}
    if shard:
        from .dist import score_nni_sharded
        res = score_nni_sharded(solver._engine, cu, ck, d0, solver.params.nu, solver.params.phi, group=group, device=device,
                                dmin=solver.dmin, dmax=solver.dmax, **kw)
    else:
        res = solver._engine.score_nni_batch(cu, ck, d0, solver.params.nu, solver.params.phi, solver.dmin, solver.dmax, **kw)
    return Neighbourhood(cu, ck, -res["nllh"], res["d5"], res["nodes"], res["iters"], d0, d_raw)


def apply_move(solver: GPU_EM_solver, nb: Neighbourhood, i: int) -> None:
    """Make candidate ``i`` the solver's current tree: the swap of ``apply_nni`` (``:281-287``) on the
    solver's own node objects and, in step, on the flat arrays -- node ids do not change under an NNI, so
    nothing is re-flattened -- the five optimised lengths, and ``pmm_set_tree`` on the live handle (the
    character matrix stays on the device)."""
    forest = solver._forest
    u_id, k = int(nb.cand_u[i]), int(nb.cand_k[i])
    u = forest.nodes[u_id]
    v = u.parent
    w = next(c for c in v.children if c is not u)
    child = u.children[k]
    for j in np.nonzero(nb.d0 != nb.d_raw)[0]:                    # the clamp of start_lengths, where it bit
        forest.nodes[int(j)].edge_length = float(nb.d0[j])
    for node_id, length in zip(nb.nodes[i], nb.d5[i]):
        forest.nodes[int(node_id)].edge_length = float(length)
    # slot swap: w takes the moved child's place under u, the child takes w's place under v -- the same
    # convention as nni.apply_nni, so children[0] / children[1] stay the arrays' left / right
    u.children[k], v.children[v.children.index(w)] = w, child
    w.parent, child.parent = u, v
    parent, left, right = apply_nni(forest.parent, forest.left, forest.right, u_id, k)
    new = forest_from_parent(parent, left, right, derive=False)      # traversal orders on demand (FlatForest.derived)
    new.nodes, new.leaf_labels = forest.nodes, forest.leaf_labels
    solver._engine.set_tree(parent, left, right)
    solver._forest = new
    solver._stats = None


class GPU_Topology_search:
    """``Topology_search`` (``laml_libs/Topology_search.py:10-216``) with the same constructor, ``search``
    signature and return value ``(best_trees, best_score, best_params)``, on one persistent solver.

    ``clamp_start`` (default False) is the one place the search departs from the reference on purpose.  The
    reference compares a candidate scored with ALL branches raised to ``2 dmin = 0.01`` (``ini_brlens`` through
    ``x2brlen``, ``ML_solver.py:195-203``, ``EM_solver.py:45-51``) against a current score obtained with
    branches as short as ``dmin = 0.005``.  Every candidate therefore carries the same penalty -- the likelihood
    lost by lengthening all short branches -- which is small on a 100-cell tree and larger than any NNI gain at
    5 000 cells: measured on BASELINE config 5 from a tree 100 NNI moves away from the generating one, not one of
    9 996 candidates was accepted (``profiles/r1_nni_search_config5.json``).  With ``clamp_start=False`` candidates
    keep the current lengths outside the five free branches, so scores are comparable with the current score;
    ``clamp_start=True`` gives the reference's literal behaviour."""

    def __init__(self, treeTopoList, solver=GPU_EM_solver, data={}, prior={}, params={}, T_cooldown=20, alpha_cooldown=0.9,
                 device: int = 0, clamp_start: bool = False, shard_candidates: bool = False, group=None, gather_device=None,
                 batch_ultra: Optional[bool] = None):
        # shard_candidates: one process per GPU (torchrun), every rank runs this same search with the same
        # ``random`` seed -- the policy is replayed identically everywhere -- and scores 1/world of each
        # neighbourhood (``dist.score_nni_sharded``)
        # batch_ultra: score the neighbourhood in one launch under ``ultra_constr`` too (the two-height M-step of
        # pmm_nni_ultra.cuh).  On by default since it has been verified and timed on a B200 (tests/test_zzz_gpu_nni_ultra.py;
        # BASELINE config 5: 9 996 constrained candidates in 7.0 ms on one GPU, 2.5 ms on eight,
        # profiles/r2j_config5_nni_n1.json / r2j_config5_nni_n8.json); PMM_NNI_ULTRA_BATCH=0 or batch_ultra=False restores
        # the reference's candidate-by-candidate sequence
        import os
        self.batch_ultra = bool(int(os.environ.get("PMM_NNI_ULTRA_BATCH", "1"))) if batch_ultra is None else bool(batch_ultra)
        self.clamp_start = clamp_start
        self.shard_candidates, self.group, self.gather_device = shard_candidates, group, gather_device
        self.treeTopoList = list(treeTopoList)
        self.solver = solver
        self.data, self.prior, self.params = data, prior, dict(params)
        self.device = device
        self.T_cooldown, self.alpha_cooldown = T_cooldown, alpha_cooldown
        self.b = 1 / (1 - 1 / self.alpha_cooldown ** self.T_cooldown)          # :29-30
        self.a = -self.b / self.alpha_cooldown ** self.T_cooldown
        self._solver: Optional[GPU_EM_solver] = None
        self.trees_scored = 0
        from .tree import read_tree_newick
        self.has_polytomy = False                                              # :18-24 (read by runner.py:301)
        for topo in self.treeTopoList:
            t = read_tree_newick(topo)
            t.suppress_unifurcations()
            if any(len(nd.children) > 2 for nd in t.traverse_preorder()):
                self.has_polytomy = True

    # ------------------------------------------------------------------ plumbing
    def get_solver(self) -> GPU_EM_solver:
        if self.device:
            return self.solver(self.treeTopoList, self.data, self.prior, self.params, device=self.device)
        return self.solver(self.treeTopoList, self.data, self.prior, self.params)

    def update_from_solver(self, mySolver):                                    # :43-48 (called by runner.py:290)
        self.treeTopoList = mySolver.get_tree_newick()
        self.params = mySolver.get_params()

    def _open(self, resolve_polytomies: bool, eps_len: float = 1e-3) -> GPU_EM_solver:
        """the search's solver for ``self.treeTopoList``; polytomies are resolved and the inserted nodes marked
        BEFORE the solver sees the trees (``__mark_polytomies__``, ``:51-66``), so that they are ordinary,
        optimisable branches of length ``eps_len`` and the only ones ``only_marked`` moves"""
        from .tree import read_tree_newick
        trees = []
        self.has_polytomy = False
        for topo in self.treeTopoList:
            t = read_tree_newick(topo)
            t.suppress_unifurcations()
            trees.append(t)
        for t in trees:
            for nd in t.traverse_preorder():
                nd.mark = False
                if len(nd.children) > 2:
                    self.has_polytomy = True
            if resolve_polytomies:
                t.resolve_polytomies()
                for nd in t.traverse_preorder():
                    if not hasattr(nd, 'mark'):
                        nd.mark = True
                        nd.edge_length = eps_len
        if self._solver is not None:
            self._solver.close()
        self.treeTopoList = [t.newick() for t in trees]
        s = self.get_solver()
        if not isinstance(s, GPU_EM_solver):
            raise TypeError("GPU_Topology_search drives GPU_EM_solver (batched scoring needs its engine handle)")
        # carry the marks over to the solver's own node objects (same newick, same traversal order)
        for t_mine, t_solver in zip(trees, s.trees):
            mine, theirs = list(t_mine.traverse_preorder()), list(t_solver.traverse_preorder())
            for a, b in zip(mine, theirs if len(mine) == len(theirs) else []):
                b.mark = a.mark
        self._solver = s
        return s

    def close(self):
        if self._solver is not None:
            self._solver.close()
            self._solver = None

    def _sync_from_solver(self):
        self.treeTopoList = self._solver.get_tree_newick()
        self.params = self._solver.get_params()

    def __accept_proposal__(self, curr_score, new_score, t):                   # :67-80
        if new_score > curr_score:
            return True
        T = max(1e-12, self.a * self.alpha_cooldown ** t + self.b)
        p = min(exp((new_score - curr_score - 1e-12) / T), 1)
        return random() < p

    def _score_all(self, strategy) -> float:
        """all numerical parameters on the current topology (``__search_one__`` ``:176-181``, ``search``
        ``:152-158``); ``optimize`` re-reads its best trees, so marks are re-attached by traversal order"""
        s = self._solver
        marks = [[getattr(nd, "mark", False) for nd in t.traverse_preorder()] for t in s.trees]
        st = deepcopy(strategy)
        st['fixed_brlen'] = None
        st['nodes_to_recompute'] = None
        score, status = s.score_tree(strategy=st)
        for t, ms in zip(s.trees, marks):
            nodes = list(t.traverse_preorder())
            if len(nodes) == len(ms):                         # nothing was collapsed
                for nd, m in zip(nodes, ms):
                    nd.mark = m
        self.trees_scored += 1
        return score

    # ------------------------------------------------------------------ the search
    def single_nni(self, curr_score, nni_iter, strategy, only_marked=False, verbose=False, sync=True):
        """``single_nni`` + ``apply_nni`` (``:221-322``): returns ``(score, n_attempts, took)``."""
        s = self._solver
        ultra = bool(strategy.get('ultra_constr', False))
        batch = self._batched(strategy) and not (ultra and not self._can_batch_ultra())
        if batch and ultra:
            # candidates the constraint leaves no room for go through score_tree, which re-reads the trees from
            # newick: start from node ids that ARE those of the current newick, so that the neighbourhood's ids
            # survive such a round trip (ids drift from the traversal order with every array-level move)
            marks0 = [[getattr(nd, "mark", False) for nd in t.traverse_preorder()] for t in s.trees]
            self._restore(s.get_tree_newick(), (s.params.nu, s.params.phi), marks0)
        forest = s._ensure()
        branches = np.nonzero((forest.left >= 0) & (forest.parent >= 0))[0]
        if only_marked:
            branches = np.array([i for i in branches if getattr(forest.nodes[i], "mark", False)], np.int64)
        if s.num_polytomy_mark and len(branches):
            # solver-inserted (never optimised) branches take no part in a move: see score_neighbourhood
            pm = np.array([bool(getattr(nd, "polytomy_mark", False)) for nd in forest.nodes], bool)
            v = forest.parent[branches]
            w = np.where(forest.left[v] == branches, forest.right[v], forest.left[v])
            branches = branches[~(pm[branches] | pm[v] | pm[w] | pm[forest.left[branches]] | pm[forest.right[branches]])]
        # the reference shuffles the branch list with ``random.shuffle`` (:231); a permutation drawn from a numpy
        # generator seeded by ``random`` is the same policy (and stays reproducible under ``random.seed``) without
        # a Python-level pass over 10^4 branches per iteration
        branches = np.random.default_rng(getrandbits(63)).permutation(branches).tolist()
        if batch:
            nb = score_neighbourhood(s, only_marked=only_marked, clamp=self.clamp_start, shard=self.shard_candidates,
                                     group=self.group, device=self.gather_device, ultra=ultra)
            self.trees_scored += len(nb.score)
            first = nb.first_of(forest.n_nodes)
            n_attempts = 0
            seq = None                                # state of the one-at-a-time path, set up on first use
            while branches:
                u = branches.pop()
                moves = [0, 1]
                shuffle(moves)
                n_attempts += 2
                for k in moves:
                    i = int(first[u]) + k
                    new_score = float(nb.score[i])
                    if new_score != new_score:
                        # the constrained local M-step had no room (or failed): the reference's sequence for such a
                        # candidate -- local attempt, then every branch free (:301-306) -- on the live solver
                        if seq is None:
                            seq = (s.get_tree_newick(), (s.params.nu, s.params.phi),
                                   [[getattr(nd, "mark", False) for nd in t.traverse_preorder()] for t in s.trees])
                        new_score, took = self._sequential_candidate(u, k, strategy, curr_score, nni_iter, *seq)
                        if took:
                            return new_score, n_attempts, True
                        s._ensure()                   # node ids are those of the same newick: nb stays valid
                        continue
                    if self.__accept_proposal__(curr_score, new_score, nni_iter):
                        apply_move(s, nb, i)
                        if sync:                      # ``sync=False``: the caller tracks snapshots (__search_one__)
                            self._sync_from_solver()
                        return new_score, n_attempts, True
            return curr_score, n_attempts, False
        # One candidate at a time -- the reference's own call sequence (:246-306) on the live solver and handle --
        # for what the batch kernel does not cover: local_brlen_opt = False (every parameter re-optimised per
        # candidate) and, unless ``batch_ultra``, the ultrametric equalities.
        before, params0 = s.get_tree_newick(), (s.params.nu, s.params.phi)
        marks = [[getattr(nd, "mark", False) for nd in t.traverse_preorder()] for t in s.trees]
        n_attempts = 0
        while branches:
            u_id = branches.pop()
            moves = [0, 1]
            shuffle(moves)
            n_attempts += 2
            for k in moves:
                new_score, took = self._sequential_candidate(u_id, k, strategy, curr_score, nni_iter, before, params0, marks)
                if took:
                    return new_score, n_attempts, True
        return curr_score, n_attempts, False

    def _sequential_candidate(self, u_id, k, strategy, curr_score, nni_iter, before, params0, marks):
        """``apply_nni`` (``:242-322``) for one candidate on the live solver: ``(new_score, accepted)``.  ``score_tree``
        replaces the solver's node objects; a rejected candidate is undone by re-reading the trees it started from
        (same newick => same node ids after flattening)."""
        s = self._solver
        local = bool(strategy.get('local_brlen_opt', True))
        forest = s._ensure()
        u = forest.nodes[u_id]
        v = u.parent
        w = next(c for c in v.children if c is not u)
        child = u.children[k]
        st = deepcopy(strategy)
        st['fixed_brlen'] = None
        st['nodes_to_recompute'] = None
        if local:                                                   # :253-277
            st['fixed_nu'], st['fixed_phi'] = s.params.nu, s.params.phi
            free = {id(x) for x in (u.children + v.children + [v])}
            fixed_brlen = []
            for tree in s.trees:
                fb = {}
                for nd in tree.traverse_postorder():
                    nd.anchors = (nd.label, nd.label) if not nd.children else \
                        (nd.children[0].anchors[0], nd.children[-1].anchors[0])
                    if id(nd) not in free:
                        fb[nd.anchors] = nd.edge_length
                fixed_brlen.append(fb)
            st['fixed_brlen'] = fixed_brlen
        u.remove_child(child); v.add_child(child)
        v.remove_child(w); u.add_child(w)
        s._forest = None
        marks_moved = [[getattr(nd, "mark", False) for nd in t.traverse_preorder()] for t in s.trees]
        new_score, status = s.score_tree(strategy=st)
        self.trees_scored += 1
        if local and status != "optimal":                           # :301-306: retry with everything free
            self._restore(before, params0, marks)
            forest = s._ensure()
            u = forest.nodes[u_id]; v = u.parent
            w = next(c for c in v.children if c is not u); child = u.children[k]
            u.remove_child(child); v.add_child(child)
            v.remove_child(w); u.add_child(w)
            s._forest = None
            st['fixed_brlen'] = None
            new_score, status = s.score_tree(strategy=st)
            self.trees_scored += 1
        if new_score is not None and self.__accept_proposal__(curr_score, new_score, nni_iter):
            for t, ms in zip(s.trees, marks_moved):                 # score_tree re-read the moved trees: marks by position
                nodes = list(t.traverse_preorder())
                for nd, m in zip(nodes, ms if len(ms) == len(nodes) else [False] * len(nodes)):
                    nd.mark = m
            self._sync_from_solver()
            return new_score, True
        self._restore(before, params0, marks)
        return new_score, False

    def _can_batch_ultra(self) -> bool:
        """the batch kernel's constrained M-step treats everything outside the five free branches as constants of
        an ultrametric tree: usable when asked for and the current lengths (as the candidates will see them) are"""
        if not self.batch_ultra:
            return False
        s = self._solver
        forest = s._ensure()
        d = s._d()
        if self.clamp_start:
            d = np.maximum(d, 2 * s.dmin)
        return is_ultrametric(forest, d, tol=5e-10)            # inside the library's own 1e-9 whatever the summation order

    @staticmethod
    def _batched(strategy) -> bool:
        return bool(strategy.get('local_brlen_opt', True))

    def _restore(self, newicks, params, marks):
        from .tree import read_tree_newick
        s = self._solver
        s.trees = [read_tree_newick(t) for t in newicks]
        if s.num_polytomy_mark:
            # the newicks may carry the branches the SOLVER inserted to binarise polytomies (never optimised, length
            # dmin / 100): contract them as ``optimize`` does (ML_solver.py:317) -- the solver re-inserts and re-marks
            # them when it binds the trees, exactly what a fresh reference solver would do with the polytomous tree
            for t in s.trees:
                t.collapse_short_branches(s.dmin * 0.01)
        for j, t in enumerate(s.trees):
            nodes = list(t.traverse_preorder())
            ms = marks[j] if marks is not None and len(marks[j]) == len(nodes) else [False] * len(nodes)
            for nd, m in zip(nodes, ms):
                nd.mark, nd.mark_fixed, nd.polytomy_mark = m, False, False
        s.params.nu, s.params.phi = params
        s._forest = None
        s._stats = None

    # -- snapshots: during a local-move search the trees are only arrays + lengths (node ids are stable under NNI);
    #    newick strings are produced when somebody needs them (checkpoints, the end), not per accepted move
    def _snapshot(self):
        s = self._solver
        f = s._ensure()
        return (f.parent.copy(), f.left.copy(), f.right.copy(), s._d(), (s.params.nu, s.params.phi))

    def _load(self, snap):
        s = self._solver
        parent, left, right, d, (nu, phi) = snap
        nodes = s._forest.nodes
        for i, nd in enumerate(nodes):
            nd.edge_length = float(d[i])
            if left[i] >= 0:
                nd.children = [nodes[left[i]], nodes[right[i]]]
            nd.parent = nodes[parent[i]] if parent[i] >= 0 else None
        new = forest_from_parent(parent, left, right, derive=False)
        new.nodes, new.leaf_labels = nodes, s._forest.leaf_labels
        s._engine.set_tree(parent, left, right)
        s._forest, s._stats = new, None
        s.params.nu, s.params.phi = nu, phi

    def _materialise(self, snap):
        """(newick strings, params) of a snapshot; the solver is left where it was"""
        here = self._snapshot()
        self._load(snap)
        out = (self._solver.get_tree_newick(), self._solver.get_params())
        self._load(here)
        return out

    def __search_one__(self, strategy, maxiter=100, verbose=False, only_marked=False, checkpoint_file=None):   # :174-219
        curr_score = self._score_all(strategy)
        if verbose:
            print("Initial score: " + str(curr_score), flush=True)
        self._sync_from_solver()
        # the one-at-a-time path (all of ultra_constr without batch_ultra, its infeasible candidates with) replaces
        # the node objects: no array snapshots there
        lazy = self._batched(strategy) and not strategy.get('ultra_constr', False)
        best_score, best = curr_score, (self._snapshot() if lazy else (self.treeTopoList, self.params))
        for nni_iter in range(maxiter):
            if verbose:
                print("NNI Iter:", nni_iter, flush=True)
            new_score, n_attempts, success = self.single_nni(curr_score, nni_iter, strategy, only_marked=only_marked,
                                                             sync=not lazy)
            if verbose:
                print(f"Number of trees checked: {n_attempts}", flush=True)
            if not success:
                break
            curr_score = new_score
            if curr_score > best_score:
                best_score, best = curr_score, (self._snapshot() if lazy else (self.treeTopoList, self.params))
            if verbose:
                print("Current score: " + str(curr_score), flush=True)
            if nni_iter % chkpt_freq == 0 and checkpoint_file is not None:
                best_trees, best_params = self._materialise(best) if lazy else best
                with open(checkpoint_file, "a") as fout:
                    fout.write(f"NNI Iteration: {nni_iter}\n")
                    fout.write(f"Current newick tree: {best_trees}\n")
                    fout.write(f"Current negative-llh: {best_score}\n")
                    fout.write(f"Current dropout rate: {best_params['phi']}\n")
                    fout.write(f"Current silencing rate: {best_params['nu']}\n")
        if verbose:
            print("Best score for this search: " + str(best_score), flush=True)
        best_trees, best_params = self._materialise(best) if lazy else best
        return best_trees, best_score, best_params

    def search(self, resolve_polytomies=True, maxiter=100, verbose=False, nreps=1, strategy=DEFAULT_STRATEGY,
               checkpoint_file=None):                                                                       # :108-172
        original_topos, original_params = self.treeTopoList, self.params
        best_trees, best_score, best_params = None, -float("inf"), None
        try:
            for i in range(nreps):
                if verbose:
                    print("Performing nni-search " + str(i + 1), flush=True)
                self.treeTopoList, self.params = original_topos, original_params
                self._open(resolve_polytomies)
                if strategy.get('resolve_search_only'):
                    if self.has_polytomy:
                        trees, score, params = self.__search_one__(strategy, maxiter=maxiter, verbose=verbose,
                                                                   only_marked=True, checkpoint_file=checkpoint_file)
                    else:
                        score = self._score_all(strategy)
                        self._sync_from_solver()
                        trees, params = self.treeTopoList, self.params
                else:
                    trees, score, params = self.__search_one__(strategy, maxiter=maxiter, verbose=verbose,
                                                               only_marked=False, checkpoint_file=checkpoint_file)
                # the final optimisation of all parameters on the chosen topology (:146-160)
                self.treeTopoList, self.params = trees, params
                self._restore(trees, (params['nu'], params['phi']), None)      # same handle: only the tree changes
                st = deepcopy(strategy)
                st['final_optimization'] = True
                score = self._score_all(st)
                self._sync_from_solver()
                if verbose:
                    print("Optimal score for this search: " + str(score), flush=True)
                if score > best_score:
                    best_score, best_trees, best_params = score, self.treeTopoList, self.params
        finally:
            self.close()
        self.treeTopoList, self.params = best_trees, best_params
        return best_trees, best_score, best_params
