"""Posterior annotation of the optimised trees.

The reference writes ``<prefix>_annotations.txt`` at the end of a run (``laml_libs/runner.py:367-435``):
it builds a throw-away ``EM_solver`` on the optimised trees, runs ``az_partition`` and ``Estep`` -- which
leave an ``alpha`` list, ``post0`` / ``post1`` lists and five ``S`` lists of length K on every node -- and
then walks the node objects.  That is n x K x 8 Python objects before a line is written.

Here the same file comes from arrays: the engine returns ``post0`` / ``post1`` ``[n][K]`` (``pmm_posteriors``)
and the per-branch sums ``S[n][5]`` (``pmm_estep``); the z / alpha / ? tags -- the one thing the full-state
engine never needs, ``ML_solver.py:125-154`` -- are integer arrays built level by level on the host
(:func:`az_tags`), the top-down filling of ``'?'`` tags (``runner.py:423-425``) is one vectorised sweep per
depth (:func:`fill_missing_tags`), and lines are formatted one node at a time.  Output is identical to the
reference's, including its quirks: ``idx`` restarts at 1 whenever a root ``I0`` is added, and the expansion
of a ``'?'`` tag uses the caller's *raw* prior dict (un-normalised, key 0 included).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

TAG_Z = 0           # 'z': the unedited state is forced / the clade is not below a single edit
TAG_MISSING = -1    # '?': nothing observed below


def az_tags(forest, leaf_codes) -> np.ndarray:
    """``ML_solver.az_partition`` (``ML_solver.py:125-154``) on arrays: ``tags[node, site]`` = ``TAG_Z``,
    ``TAG_MISSING`` or the 1-based rank ``a`` of the single edited state observed below the node.
    ``leaf_codes``: ``(n_leaves, K)`` site codes in leaf-id order (``flatten.encode_sites``); both kinds of
    missing entry (``'?'`` and the literal -1) give ``'?'`` as in the reference (``:137-140``)."""
    leaf_codes = np.asarray(leaf_codes)
    n, K = forest.n_nodes, leaf_codes.shape[1]
    tags = np.empty((n, K), np.int32)
    tags[: forest.n_leaves] = np.where(leaf_codes < 0, TAG_MISSING, leaf_codes)
    height = forest.height
    for h in range(1, int(height.max()) + 1 if n else 0):
        idx = np.nonzero(height == h)[0]
        a, b = tags[forest.left[idx]], tags[forest.right[idx]]
        z = (a == TAG_Z) | (b == TAG_Z) | ((a > 0) & (b > 0) & (a != b))          # :148-149
        tags[idx] = np.where(z, TAG_Z, np.where(a > 0, a, np.where(b > 0, b, TAG_MISSING)))
    return tags


def fill_missing_tags(forest, tags) -> np.ndarray:
    """``runner.py:423-425``: walking down from the root, a ``'?'`` tag takes the parent's tag unless that is
    ``'z'``.  The root added above a tree carries ``'z'`` (``:402``), so a tree's own root keeps its tags."""
    tags = tags.copy()
    depth = forest.depth
    for lvl in range(1, int(depth.max()) + 1 if forest.n_nodes else 0):
        idx = np.nonzero(depth == lvl)[0]
        t, pt = tags[idx], tags[forest.parent[idx]]
        tags[idx] = np.where((t == TAG_MISSING) & (pt != TAG_Z), pt, t)
    return tags


def format_posterior(p0, p_minus_1, p_alpha, alpha: str, q: dict) -> str:
    """One site entry (``runner.py:368-389``): a state alone when its rounded posterior is 1, otherwise
    ``state:prob`` items joined by ``/``; an unresolved ``'?'`` spreads ``p_alpha`` over the prior ``q``."""
    if p0 == 1:
        return '0'
    if p_minus_1 == 1:
        return '-1'
    if p_alpha == 1:
        return alpha
    items = []
    if p0 > 0:
        items.append('0:' + str(p0))
    if p_minus_1 > 0:
        items.append('-1:' + str(p_minus_1))
    if p_alpha > 0:
        if alpha == '?':
            # the reference joins this list onto the '/' it has already appended, so an empty list leaves a
            # trailing '/', and as the only item it yields '' -- kept
            items.append("/".join(str(y) + ':' + str(round(p_alpha * q[y], 3)) for y in q if round(p_alpha * q[y], 3) > 0))
        else:
            items.append(alpha + ":" + str(p_alpha))
    return "/".join(items)


def _row_text(tag_row, p0_row, p1_row, alpha_names: Sequence[Sequence[str]], Q) -> str:
    out = []
    for j, (t, a, m) in enumerate(zip(tag_row, p0_row, p1_row)):
        p0 = round(a, 2)
        if p0 == 1:
            out.append('0')
            continue
        pm = round(m, 2)
        pa = round(1 - p0 - pm, 2)
        alpha = 'z' if t == TAG_Z else ('?' if t == TAG_MISSING else alpha_names[j][t - 1])
        out.append(format_posterior(p0, pm, pa, alpha, Q[j]))
    return ",".join(out)


def write_annotations(solver, Q, out_annotate) -> List[str]:
    """Write the reference's annotation file for ``solver`` (a ``GPU_EM_solver`` at the parameters to
    annotate).  ``Q`` is the prior as the caller holds it (list of dicts; the runner passes its raw, possibly
    un-normalised one, ``runner.py:433``).  The solver's trees are left as they were.  Returns the annotated
    newick strings (labelled nodes, branch lengths = expected mutations): what the runner reads off
    ``my_solver.trees`` for its return value (``runner.py:450-451``)."""
    from .tree import Node
    forest = solver._ensure().derived()
    enc = solver._enc
    K = enc.n_sites
    d = solver._d()
    est = solver._engine.estep(d, solver.params.nu, solver.params.phi)
    post0, post1 = solver._engine.posteriors(d, solver.params.nu, solver.params.phi)
    S = est["S"]
    expected_mutations = S[:, 1] + S[:, 2] + S[:, 4]                          # runner.py:414
    tags = fill_missing_tags(forest, az_tags(forest, enc.codes[solver._leaf_row_bound]))
    alpha_names: List[List[str]] = [[str(int(s)) for s in enc.states[enc.qptr[j]:enc.qptr[j + 1]]] for j in range(K)]
    node_id = {id(nd): i for i, nd in enumerate(forest.nodes)}
    with np.errstate(under="ignore"):
        P0, P1 = np.exp(post0), np.exp(post1)

    saved = [(nd, nd.label, nd.edge_length) for nd in forest.nodes]
    old_roots = [tree.root for tree in solver.trees]
    idx = 0
    annotated: List[str] = []
    try:
        with open(out_annotate, 'w') as fout:
            for tree in solver.trees:
                added = None
                if len(tree.root.children) > 1:                                # runner.py:397-408
                    added = Node()
                    added.label = 'I0'
                    idx = 1
                    added.add_child(tree.root)
                    tree.root = added
                all_labels = set()
                for node in tree.traverse_preorder():                         # runner.py:409-417
                    if node.is_root():
                        continue
                    if node.label is None or node.label in all_labels:
                        node.label = 'I' + str(idx)
                        idx += 1
                    all_labels.add(node.label)
                    node.edge_length = round(float(expected_mutations[node_id[id(node)]]), 3)
                annotated.append(tree.newick())
                fout.write(annotated[-1] + "\n")
                for node in tree.traverse_preorder():                         # runner.py:419-435
                    if node is added:
                        fout.write(node.label + "," + ",".join(['0'] * K) + "\n")
                        continue
                    i = node_id[id(node)]
                    fout.write(node.label + "," + _row_text(tags[i].tolist(), P0[i].tolist(), P1[i].tolist(),
                                                             alpha_names, Q) + "\n")
    finally:
        for tree, root in zip(solver.trees, old_roots):
            if tree.root is not root:
                tree.root.remove_child(root)
                root.parent = None
                tree.root = root
        for nd, label, length in saved:
            nd.label, nd.edge_length = label, length
    return annotated


def annotate(opt_trees, msa, Q, opt_params, out_annotate, device: int = 0) -> List[str]:
    """The whole annotation step of the runner (``runner.py:391-435``) in one call: a throw-away solver on
    the optimised trees at the optimised parameters, one E-step on the GPU, the file.  Returns the annotated
    newick strings (see :func:`write_annotations`)."""
    from .solver import GPU_EM_solver
    solver = GPU_EM_solver(opt_trees, {'charMtrx': msa}, {'Q': Q}, {'phi': opt_params['phi'], 'nu': opt_params['nu']},
                           device=device)
    try:
        return write_annotations(solver, Q, out_annotate)
    finally:
        solver.close()
