"""The call the reference makes to score ONE NNI candidate (``Topology_search.apply_nni``,
``laml_libs/Topology_search.py:242-300``), restated for the parity tests of the batched scorer: anchors of
every branch outside the five free ones, their current lengths as ``fixed_brlen``, nu and phi fixed, a NEW
solver built from the newick of the re-wired tree, ``score_tree``."""
from laml_b200.solver import DEFAULT_STRATEGY, GPU_EM_solver
from laml_b200.tree import read_tree_newick


def score_candidate_like_the_reference(solver, u_id, k, case, strategy_extra=None):
    forest = solver._ensure()
    trees = [read_tree_newick(t) for t in solver.get_tree_newick()]          # treeList_obj
    # locate u in the copy by traversal position
    pos = None
    for ti, t in enumerate(solver.trees):
        for j, nd in enumerate(t.traverse_preorder()):
            if nd is forest.nodes[u_id]:
                pos = (ti, j)
    u = list(trees[pos[0]].traverse_preorder())[pos[1]]
    v = u.parent
    w = next(c for c in v.children if c is not u)
    u_child = u.children[k]
    free = set(u.child_nodes() + v.child_nodes() + [v])                     # :258
    fixed_brlen = [{} for _ in trees]
    for ti, t in enumerate(trees):                                          # :261-277
        for nd in t.traverse_postorder():
            if nd.is_leaf():
                nd.anchors = (nd.label, nd.label)
            else:
                C = nd.child_nodes()
                nd.anchors = (C[0].anchors[0], C[-1].anchors[0])
            if nd not in free:
                fixed_brlen[ti][nd.anchors] = nd.edge_length
    u.remove_child(u_child); v.add_child(u_child)                           # :281-287
    v.remove_child(w); u.add_child(w)
    st = dict(DEFAULT_STRATEGY)
    st.update(fixed_brlen=fixed_brlen, fixed_nu=solver.params.nu, fixed_phi=solver.params.phi)
    st.update(strategy_extra or {})
    new = GPU_EM_solver([t.newick() for t in trees], {'charMtrx': case.msa}, {'Q': case.Q}, solver.get_params())
    score, status = new.score_tree(strategy=st)
    new.close()
    return score, status
