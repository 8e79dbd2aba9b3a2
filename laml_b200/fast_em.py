"""The array-level boundary of the reference's ``fast_em`` engine on the B200 engine.

``fast_em/laml.py:196-276`` (``EMOptimizer``) is the reference's own array interface to a PMM EM: priors
as a dense ``(num_characters, max_alphabet_size)`` array, the character matrix as ``(num_leaves,
num_characters)`` integers (0 unedited, ``1..A`` the recoded states, -1 unknown, ``fast_em/phylogeny.py:
24-35``), the tree as a ``networkx.DiGraph`` over node ids ``0..2N-2`` with the leaves first (``fastEM_solver.
parse_tree``, ``laml_libs/fastEM_solver.py:46-146``), branch lengths and a branch mask as vectors.  There it
feeds a float32 JAX program whose M-step is 100 gradient steps; here the same constructor and ``optimize``
signature drive ``pmm_em`` (fp64, per-branch Newton M-step), so ``fastEM_solver`` can switch engines by
changing one import.

Differences a caller sees: results are float64 numpy values; convergence is EM_solver's relative 1e-8
(``conv_eps=`` to change; fast_em stops at 1e-5); branch lengths are kept in EM_solver's box
``[dmin, dmax] = [0.005, 10]``.  Prior rows are used as given, zero-padded entries are states of prior 0
(fast_em takes ``log(mutation_priors)`` of the same rows, ``fast_em/calculations.py:150``) -- pass rows
normalised as ``ML_solver.py:70-74`` does to reproduce ``EM_solver``.
"""
from __future__ import annotations

import numpy as np

from .capi import Engine


class EMOptimizer:
    def __init__(self, mutation_priors, character_matrix, verbose=False, device: int = 0):
        self.mutation_priors = np.asarray(mutation_priors, np.float64)          # (num_characters, max_alphabet_size)
        self.character_matrix = np.asarray(character_matrix, np.int32)          # (num_leaves, num_characters)
        self.num_leaves = int(self.character_matrix.shape[0])
        self.alphabet_size = int(self.mutation_priors.shape[1])
        self.num_characters = int(self.character_matrix.shape[1])
        self.verbose = verbose
        if self.mutation_priors.shape[0] != self.num_characters:
            raise ValueError("mutation_priors has %d rows, character_matrix %d columns"
                             % (self.mutation_priors.shape[0], self.num_characters))
        if self.character_matrix.min(initial=0) < -1 or self.character_matrix.max(initial=0) > self.alphabet_size:
            raise ValueError("character states must be -1 (unknown), 0, or 1..%d" % self.alphabet_size)
        K, A = self.num_characters, self.alphabet_size
        qptr = (np.arange(K + 1, dtype=np.int64) * A).astype(np.int32)
        self._engine = Engine(self.num_leaves, qptr, self.mutation_priors.reshape(-1), self.character_matrix, device=device)
        self._R = float((self.character_matrix != -1).sum())                    # EM_solver.py:322

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    @staticmethod
    def _flatten(tree, n_leaves):
        n = 2 * n_leaves - 1
        if tree.number_of_nodes() != n:
            raise ValueError("tree must be binary over node ids 0..%d" % (n - 1))
        parent = np.full(n, -1, np.int32); left = np.full(n, -1, np.int32); right = np.full(n, -1, np.int32)
        for v in range(n):
            succ = list(tree.successors(v))
            if len(succ) == 0:
                if v >= n_leaves:
                    raise ValueError("leaves must be the node ids 0..%d" % (n_leaves - 1))
                continue
            if len(succ) != 2 or v < n_leaves:
                raise ValueError("node %d: internal nodes have two children and ids >= num_leaves" % v)
            left[v], right[v] = int(succ[0]), int(succ[1])
            parent[left[v]] = parent[right[v]] = v
        return parent, left, right

    def optimize(self, tree, branch_lengths, nu: float, phi: float, branch_mask, fixed_nu: bool = False,
                 fixed_phi: bool = False, conv_eps: float = 1e-8, max_iter: int = 1000, dmin: float = 0.005,
                 dmax: float = 10.0):
        """``EMOptimizer.optimize`` (``fast_em/laml.py:207-276``): EM from the given point; ``branch_mask[i]``
        true = branch ``i`` is optimised.  Returns the same dict: ``nllh``, ``branch_lengths``, ``nu``,
        ``phi``, ``em_iterations``."""
        parent, left, right = self._flatten(tree, self.num_leaves)
        self._engine.set_tree(parent, left, right)
        d0 = np.asarray(branch_lengths, np.float64)
        mask = np.asarray(branch_mask).astype(bool).astype(np.uint8)
        res = self._engine.em(d0, float(nu), float(phi), self._R, self.num_leaves, free_mask=mask,
                              optimize_nu=not fixed_nu, optimize_phi=not fixed_phi, dmin=dmin, dmax=dmax,
                              conv_eps=conv_eps, max_iter=max_iter)
        if self.verbose:
            print("EM finished after %d iterations, nllh %.9f" % (res["iters"], res["nllh"]))
        return {"nllh": res["nllh"], "branch_lengths": res["d"], "nu": res["nu"], "phi": res["phi"],
                "em_iterations": res["iters"]}
