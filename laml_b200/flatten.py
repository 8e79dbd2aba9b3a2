"""Host-side boundary: trees and character matrices -> flat arrays.

The reference keeps trees as linked node objects and the character matrix as a
``dict[cell] -> list`` (``laml_libs/ML_solver.py:19-86``) and walks them with
Python loops.  The engine takes *arrays*; this module is the only place where
the two meet.  The array conventions follow the reference's own array engine
precedent (leaves first, then internal nodes: ``laml_libs/fastEM_solver.py:82-134``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np

CODE_MISSING = -1   # '?'  (dropout or silenced, unknown which)
CODE_SILENCED = -2  # literal -1 in the matrix: known silenced (ML_solver.py:139-140)


@dataclass
class FlatForest:
    """A forest of rooted binary trees as int32 arrays.

    ids: leaves ``0..n_leaves-1`` then internal nodes; ``order`` is a
    children-before-parents sequence (the reference's postorder,
    ``EM_solver.py:97``); ``parent``/``left``/``right`` use -1 for "none".
    """
    parent: np.ndarray
    left: np.ndarray
    right: np.ndarray
    order: np.ndarray
    depth: np.ndarray          # root = 0
    height: np.ndarray         # leaf = 0
    n_leaves: int
    nodes: list = field(default_factory=list, repr=False)   # node objects by id (may be empty)
    leaf_labels: List[str] = field(default_factory=list, repr=False)

    @property
    def n_nodes(self) -> int:
        return int(self.parent.shape[0])

    @property
    def roots(self) -> np.ndarray:
        return np.nonzero(self.parent < 0)[0].astype(np.int32)

    def edge_lengths(self) -> np.ndarray:
        return np.array([float(nd.edge_length) if nd.edge_length is not None else np.nan
                         for nd in self.nodes], dtype=np.float64)

    def derived(self) -> "FlatForest":
        """this forest with ``order`` / ``depth`` / ``height`` present (see ``forest_from_parent(derive=False)``)"""
        if self.order is None:
            full = forest_from_parent(self.parent, self.left, self.right)
            self.order, self.depth, self.height = full.order, full.depth, full.height
        return self


def flatten_trees(trees: Sequence, leaf_key=None) -> FlatForest:
    """Flatten tree objects (anything with ``traverse_postorder`` and nodes
    carrying ``children``/``parent``) into a :class:`FlatForest`.  All trees
    must be binary (the solver binarises first, ``EM_solver.py:26-43``).
    ``leaf_key`` (label -> sortable) fixes the order of the leaf ids inside each
    tree independently of the topology, so an engine that already holds the
    character matrix in that order only needs ``set_tree`` for a new topology."""
    leaves, internals = [], []
    for tree in trees:
        tree_leaves = []
        for nd in tree.traverse_postorder():
            k = len(nd.children)
            if k == 0:
                tree_leaves.append(nd)
            elif k == 2:
                internals.append(nd)
            else:
                raise ValueError("PMM engine needs binary trees; found a node with %d children" % k)
        if leaf_key is not None:
            tree_leaves.sort(key=lambda nd: leaf_key[nd.label])
        leaves.extend(tree_leaves)
    nodes = leaves + internals
    ids = {id(nd): i for i, nd in enumerate(nodes)}
    n = len(nodes)
    parent = np.full(n, -1, np.int32)
    left = np.full(n, -1, np.int32)
    right = np.full(n, -1, np.int32)
    for i, nd in enumerate(nodes):
        if nd.children:
            left[i] = ids[id(nd.children[0])]
            right[i] = ids[id(nd.children[1])]
            parent[left[i]] = i
            parent[right[i]] = i
    order = []
    for tree in trees:
        order.extend(ids[id(nd)] for nd in tree.traverse_postorder())
    order = np.asarray(order, np.int32)
    depth, height = depth_and_height(parent, left, right, order)
    return FlatForest(parent, left, right, order, depth, height, len(leaves), nodes,
                      [str(nd.label) for nd in leaves])


def depth_and_height(parent, left, right, order):
    n = parent.shape[0]
    depth = np.zeros(n, np.int32)
    height = np.zeros(n, np.int32)
    for v in order:                      # children first
        if left[v] >= 0:
            height[v] = 1 + max(height[left[v]], height[right[v]])
    for v in order[::-1]:                # parents first
        if parent[v] >= 0:
            depth[v] = depth[parent[v]] + 1
    return depth, height


def forest_from_parent(parent: np.ndarray, left: np.ndarray, right: np.ndarray, derive: bool = True) -> FlatForest:
    """Build a :class:`FlatForest` straight from arrays (synthetic workloads):
    leaves must already be ids ``0..n_leaves-1``.  ``derive=False`` leaves ``order`` / ``depth`` / ``height``
    unset (None) for callers that only need the topology -- an NNI move on a live solver, where deriving them
    would cost as much as scoring the move; :meth:`FlatForest.derived` fills them in on demand."""
    parent = np.ascontiguousarray(parent, np.int32)
    left = np.ascontiguousarray(left, np.int32)
    right = np.ascontiguousarray(right, np.int32)
    n = parent.shape[0]
    is_leaf = left < 0
    n_leaves = int(is_leaf.sum())
    if not is_leaf[:n_leaves].all():
        raise ValueError("leaves must come first")
    if not derive:
        return FlatForest(parent, left, right, None, None, None, n_leaves)
    # iterative postorder via depth: sort by decreasing depth is children-before-parents
    depth = np.zeros(n, np.int32)
    roots = np.nonzero(parent < 0)[0]
    frontier = roots
    while frontier.size:
        kids = np.concatenate([left[frontier], right[frontier]])
        kids = kids[kids >= 0]
        depth[kids] = depth[parent[kids]] + 1
        frontier = kids
    order = np.argsort(-depth, kind="stable").astype(np.int32)
    height = np.zeros(n, np.int32)
    for lvl in range(int(depth.max()), -1, -1):
        idx = np.nonzero((depth == lvl) & ~is_leaf)[0]
        if idx.size:
            height[idx] = 1 + np.maximum(height[left[idx]], height[right[idx]])
    return FlatForest(parent, left, right, order, depth, height, n_leaves)


@dataclass
class SiteEncoding:
    """Per-site alphabets and the integer-coded character matrix.

    ``qptr[s]..qptr[s+1]`` indexes the edited states of site ``s`` in ``qval``
    (normalised prior, ``ML_solver.py:70-74``) and ``states`` (original state
    ids, ascending).  ``codes[row, s]``: 0 unedited, ``a>=1`` the 1-based rank of
    the observed state in the site's alphabet, -1 ``'?'``, -2 known silenced.
    """
    qptr: np.ndarray
    qval: np.ndarray
    states: np.ndarray
    codes: np.ndarray
    cells: List[str]

    @property
    def n_sites(self) -> int:
        return int(self.qptr.shape[0] - 1)


def encode_sites(charMtrx: Dict[str, list], Q: Sequence[dict]) -> SiteEncoding:
    """Integer-code a reference-style character matrix against the prior.

    The prior file defines the alphabet (``ML_solver.py:34-46`` keeps unobserved
    states); key 0, when present, takes part in the normalisation only
    (``ML_solver.py:70-74``) and is not an edited state."""
    cells = list(charMtrx.keys())
    K = len(charMtrx[cells[0]]) if cells else 0
    if len(Q) < K:
        raise ValueError("prior has %d sites, character matrix has %d" % (len(Q), K))
    qptr = np.zeros(K + 1, np.int32)
    qval: List[float] = []
    states: List[int] = []
    rank: List[Dict[int, int]] = []
    for s in range(K):
        qs = Q[s]
        tot = sum(qs[x] for x in qs)
        keys = sorted(k for k in qs if k != 0)
        rank.append({k: i + 1 for i, k in enumerate(keys)})
        states.extend(keys)
        qval.extend(qs[k] / tot for k in keys)
        qptr[s + 1] = len(states)
    codes = np.empty((len(cells), K), np.int32)
    for r, cell in enumerate(cells):
        row = charMtrx[cell]
        for s in range(K):
            x = row[s]
            if x == 0:
                codes[r, s] = 0
            elif x == "?":
                codes[r, s] = CODE_MISSING
            elif x == -1:
                codes[r, s] = CODE_SILENCED
            else:
                try:
                    codes[r, s] = rank[s][x]
                except KeyError:
                    raise KeyError("state %r at site %d of cell %r is not in the prior" % (x, s, cell))
    return SiteEncoding(qptr, np.asarray(qval, np.float64), np.asarray(states, np.int64), codes, cells)
