"""Array-level NNI moves (``laml_libs/Topology_search.py:221-322``): enumerate the candidates
around every internal non-root node and apply one to flat tree arrays.  The search policy itself
(shuffling, simulated annealing acceptance, checkpoints; ``Topology_search.py:67-216``) stays with
the reference's search layer; this module only provides what batched scoring needs."""
from __future__ import annotations

import numpy as np


def nni_candidates(parent, left, n_leaves):
    """All moves ``(u, k)``: swap child ``k`` (0 = left, 1 = right) of the internal, non-root node
    ``u`` with ``u``'s sibling -- the two moves ``apply_nni`` tries per branch."""
    parent = np.asarray(parent)
    us = [u for u in range(n_leaves, len(parent)) if parent[u] >= 0]
    cu = np.repeat(np.asarray(us, np.int32), 2)
    ck = np.tile(np.array([0, 1], np.int32), len(us))
    return cu, ck


def apply_nni(parent, left, right, u, k):
    """Topology after the move ``(u, k)``; node ids (and therefore branch-length indices) are kept:
    the moved child and the former sibling keep their own branch lengths, as the reference's node
    objects do (``Topology_search.py:289-296``)."""
    parent, left, right = (np.array(a, np.int32, copy=True) for a in (parent, left, right))
    v = parent[u]
    if v < 0 or left[u] < 0:
        raise ValueError("u must be an internal, non-root node")
    w = right[v] if left[v] == u else left[v]
    ck = right[u] if k else left[u]
    if k:
        right[u] = w
    else:
        left[u] = w
    if left[v] == w:
        left[v] = ck
    else:
        right[v] = ck
    parent[w], parent[ck] = u, v
    return parent, left, right
