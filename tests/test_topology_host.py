"""NNI search on the batched scorer on CPU, with the engine swapped for the
oracle-backed stand-in (tests/oracle_engine.py -- test infrastructure only; its ``score_nni_batch`` is a
full-tree EM per candidate, i.e. the definition the GPU scorer is tested against in tests/test_gpu_nni.py).
What is checked here is the HOST side: candidate enumeration, the call the reference's ``apply_nni`` makes
(new solver from the re-wired newick + ``score_tree`` with ``fixed_brlen`` anchors, ``Topology_search.py:
253-300``) giving the same score as the batch entry, an accepted move on the live handle, and the
reference's own search tests (laml_unit_tests/unit_tests_TopoSearch.py) against their brute-force optima."""
from copy import deepcopy

import numpy as np
import pytest

import laml_b200.solver as solver_mod
from helpers import load_npz_case
from laml_b200.solver import DEFAULT_STRATEGY, GPU_EM_solver
from laml_b200.topology import GPU_Topology_search, apply_move, score_neighbourhood, start_lengths
from oracle_engine import OracleEngine
from topo_reference_call import score_candidate_like_the_reference


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    OracleEngine.instances = 0


def _solver(case):
    return GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': case.phi, 'nu': case.nu})


@pytest.mark.parametrize("name", ["ref_rand_small", "ref_rand_forest"])
def test_batch_scores_equal_the_reference_call(name):
    case, z = load_npz_case(name)
    s = _solver(case)
    nb = score_neighbourhood(s)
    n_int = sum(1 for t in s.trees for nd in t.traverse_preorder() if nd.children and nd.parent is not None)
    assert len(nb.score) == 2 * n_int
    for i in range(len(nb.score)):
        ref_score, status = score_candidate_like_the_reference(s, int(nb.cand_u[i]), int(nb.cand_k[i]), case)
        assert status == "optimal"
        assert abs(ref_score - nb.score[i]) <= 1e-9 * abs(ref_score), (i, ref_score, nb.score[i])


def test_accepted_move_lives_on_the_same_handle():
    case, z = load_npz_case("ref_rand_small")
    s = _solver(case)
    nb = score_neighbourhood(s)
    i = int(np.argmax(nb.score))
    eng = s._engine
    apply_move(s, nb, i)
    assert s._engine is eng and OracleEngine.instances == 1
    assert abs(-s.negative_llh() - nb.score[i]) <= 1e-9 * abs(nb.score[i])
    fresh = GPU_EM_solver(s.get_tree_newick(), {'charMtrx': case.msa}, {'Q': case.Q}, s.get_params())
    assert abs(fresh.negative_llh() - s.negative_llh()) <= 1e-9 * abs(s.negative_llh())
    # the five optimised lengths are on the tree, every other branch sits at the clamped start
    d_now = s._d()
    lab_now = {id(nd): j for j, nd in enumerate(s._forest.nodes)}
    assert np.isfinite(d_now).all() and (d_now >= s.dmin - 1e-15).all()


def test_only_marked_restricts_the_candidates():
    case, z = load_npz_case("ref_rand_small")
    s = _solver(case)
    f = s._ensure()
    internal = [nd for nd in f.nodes if nd.children and nd.parent is not None]
    internal[1].mark = True
    nb = score_neighbourhood(s, only_marked=True)
    assert len(nb.score) == 2 and all(f.nodes[u] is internal[1] for u in nb.cand_u)
    full = score_neighbourhood(s)
    idx = full.index()
    for j in range(2):
        assert nb.score[j] == full.score[idx[(int(nb.cand_u[j]), int(nb.cand_k[j]))]]


Q5 = [{0: 0, 1: 1.0}] * 5
MSA_A = {'a': [1, 1, 0, 0, 0], 'b': [1, 1, 1, 0, 0], 'c': [0, 0, 0, 1, 0], 'd': [0, 0, 0, 1, 0]}
MSA_B = {'a': [0, 1, 1, 1, 1], 'b': [1, 0, 0, 0, 0], 'c': [1, 0, 0, 0, 0], 'd': [0, 1, 1, 1, 1]}


def _all_topologies(leaves):
    """every rooted binary topology on the leaves, by inserting one leaf at a time on every branch
    (what unit_tests_TopoSearch.py:11-49 enumerates)"""
    def insert(tree, leaf):
        # tree: nested tuples / str; returns all trees with `leaf` attached above every node
        out = [(tree, leaf)]
        if isinstance(tree, tuple):
            l, r = tree
            out += [(t, r) for t in insert(l, leaf)] + [(l, t) for t in insert(r, leaf)]
        return out
    trees = [leaves[0]]
    for leaf in leaves[1:]:
        trees = [t for base in trees for t in insert(base, leaf)]
    def nwk(t):
        return t if isinstance(t, str) else "(" + nwk(t[0]) + "," + nwk(t[1]) + ")"
    return [nwk(t) + ";" for t in trees]


def _brute_force(msa, Q, ultra_constr=False):
    best = float("inf")
    for T in _all_topologies(sorted(msa)):
        s = GPU_EM_solver([T], {'charMtrx': msa}, {'Q': Q})
        nllh, _ = s.optimize(initials=1, verbose=-1, ultra_constr=ultra_constr)
        best = min(best, nllh)
    return best


def test_topology_enumeration():
    assert len(_all_topologies(list("abcd"))) == 15 and len(set(_all_topologies(list("abcde")))) == 105


@pytest.mark.parametrize("no,msa,T0,resolve_only,nllh_bf", [
    (1, MSA_A, '((a,c),(b,d));', False, None),                    # unit_tests_TopoSearch.py:66-80   live brute force
    (3, MSA_A, '(a,b,c,d);', False, None),                        # :101-115  polytomy start
    (4, MSA_A, '((a,b),c,d);', True, None),                       # :118-134  only resolve polytomies
    (5, MSA_A, '((a,c),(b,d));', True, 11.809140958825493),       # :137-153  nothing to resolve: the optimum of T0 itself
    (6, MSA_B, '((a,b),(c,d));', False, None),                    # :156-172
    (7, MSA_B, '((a,b),c,d);', True, 10.083048489886435),         # :175-191  the reference's stored constants
    (8, MSA_B, '((c,d),a,b);', True, 10.08306053772235),          # :194-210
])
def test_reference_search_tests(no, msa, T0, resolve_only, nllh_bf):
    """the reference asserts its NNI result against a brute-force optimum to 4 places"""
    from random import seed
    seed(a=1000 + no)
    Q = [dict(q) for q in Q5]
    if nllh_bf is None:
        nllh_bf = _brute_force(msa, Q)
    ts = GPU_Topology_search([T0], GPU_EM_solver, data={'charMtrx': msa}, prior={'Q': Q}, params={'phi': 0, 'nu': 0})
    st = deepcopy(DEFAULT_STRATEGY)
    st['resolve_search_only'] = resolve_only
    best_trees, max_score, best_params = ts.search(maxiter=200, verbose=False, strategy=st, nreps=1)
    assert abs(nllh_bf + max_score) < 5e-5, (no, -max_score, nllh_bf, best_trees)          # places=4
    assert set(best_params) == {'phi', 'nu'}


@pytest.mark.parametrize("no,T0", [(9, '((a,b),(c,d));'), (10, '(a,b,c,d);')])
def test_reference_search_tests_ultrametric(no, T0):
    """unit_tests_TopoSearch.py:213-250 (tests 9 and 10): ultra_constr search against the ultrametric brute force"""
    from random import seed
    seed(a=1000 + no)
    Q = [dict(q) for q in Q5]
    nllh_bf = _brute_force(MSA_B, Q, ultra_constr=True)
    # the reference's own pre-computed brute-force value (unit_tests_TopoSearch.py:237, MOSEK M-step) -- an
    # independent pin of laml_b200/ultrametric.py, to the 4 places the reference's tests use
    assert abs(nllh_bf - 7.0063474330891236) < 5e-5
    ts = GPU_Topology_search([T0], GPU_EM_solver, data={'charMtrx': MSA_B}, prior={'Q': Q}, params={'phi': 0, 'nu': 0})
    st = deepcopy(DEFAULT_STRATEGY)
    st['ultra_constr'] = True
    best_trees, max_score, best_params = ts.search(maxiter=200, verbose=False, strategy=st, nreps=1)
    assert abs(nllh_bf + max_score) < 5e-5, (no, -max_score, nllh_bf, best_trees)


def test_search_uses_one_handle_and_improves():
    from random import seed
    seed(a=77)
    case, z = load_npz_case("ref_rand_small")
    # scramble the topology: the generating tree should score better than a random re-wiring
    s0 = _solver(case)
    start = -s0.negative_llh()
    ts = GPU_Topology_search(case.newicks, GPU_EM_solver, data={'charMtrx': case.msa}, prior={'Q': case.Q},
                             params={'phi': case.phi, 'nu': case.nu})
    OracleEngine.instances = 0
    best_trees, best_score, best_params = ts.search(maxiter=5, nreps=1)
    assert OracleEngine.instances == 1                     # one engine for the whole search
    assert best_score >= start - 1e-9
    assert ts.trees_scored >= 14


def test_best_snapshot_materialises_to_the_best_score(tmp_path):
    """__search_one__ tracks the best tree as arrays; the newick it finally returns must evaluate to the best
    score, also when annealing accepted worse moves afterwards, and checkpoints hold the same thing"""
    from random import seed
    seed(a=5)
    case, z = load_npz_case("ref_rand_ragged")
    ts = GPU_Topology_search(case.newicks, GPU_EM_solver, data={'charMtrx': case.msa}, prior={'Q': case.Q},
                             params={'phi': case.phi, 'nu': case.nu}, T_cooldown=2000, alpha_cooldown=0.999)   # hot: accepts worse moves
    ts._open(True)
    ck = tmp_path / "ckpt.txt"
    trees, score, params = ts.__search_one__(deepcopy(DEFAULT_STRATEGY), maxiter=12, checkpoint_file=str(ck))
    current = -ts._solver.negative_llh()
    fresh = GPU_EM_solver(trees, {'charMtrx': case.msa}, {'Q': case.Q}, params)
    assert abs(-fresh.negative_llh() - score) <= 1e-9 * abs(score)
    assert score >= current - 1e-9                                     # best is at least the place the walk ended
    text = ck.read_text()
    assert text.count("NNI Iteration:") == 2 and "Current newick tree: ['" in text
    # the solver still sits on the walk's last tree, node objects and arrays in step
    f = ts._solver._forest
    for i, nd in enumerate(f.nodes):
        if nd.children:
            assert f.nodes[f.left[i]] is nd.children[0] and f.nodes[f.right[i]] is nd.children[1]
            assert nd.children[0].parent is nd and nd.children[1].parent is nd
    ts.close()


def test_clamp_modes():
    """clamp=True is the reference's candidate (every length raised to 2 dmin); clamp=False keeps the other
    branches where they are.  Either way a local EM cannot end below the re-wired tree's likelihood at its
    starting lengths; and after a full EM the clamp alone costs likelihood, which is why the search compares
    un-clamped candidates (GPU_Topology_search docstring)."""
    from laml_b200.nni import apply_nni
    case, z = load_npz_case("ref_rand_ragged")
    s = _solver(case)
    s.optimize(initials=1, random_seeds=3, verbose=-1)                 # branches with no support end at dmin
    f = s._ensure()
    d_raw = s._d()
    assert (d_raw < 2 * s.dmin).sum() >= 3
    at_raw = -s.negative_llh()
    d0 = start_lengths(s)
    at_clamped = s._engine.inside(d0, s.params.nu, s.params.phi)
    assert at_clamped < at_raw - 1e-3                                  # the uniform penalty every clamped candidate carries
    probe = OracleEngine(f.n_leaves, s._enc.qptr, s._enc.qval, s._enc.codes[s._leaf_row_bound])
    for clamp, start in ((True, d0), (False, d_raw)):
        nb = score_neighbourhood(s, clamp=clamp)
        assert (nb.d0 == start).all()
        for i in range(0, len(nb.score), 7):
            p, l, r = apply_nni(f.parent, f.left, f.right, int(nb.cand_u[i]), int(nb.cand_k[i]))
            probe.set_tree(p, l, r)
            assert nb.score[i] >= probe.inside(start, s.params.nu, s.params.phi) - 1e-9
    # un-clamped candidates are comparable with the current score: the best one is not far below it, whereas
    # every clamped one sits below by at least most of the penalty
    assert score_neighbourhood(s, clamp=False).score.max() > score_neighbourhood(s, clamp=True).score.max()


def test_search_from_a_perturbed_tree_climbs():
    from random import seed
    from laml_b200.flatten import forest_from_parent
    from laml_b200.nni import apply_nni
    from laml_b200.synth import Workload, make_workload, reference_inputs
    w = make_workload(24, 40, 2, 8, 0.1, 0.1, seed=12)
    rng = np.random.default_rng(1)
    p, l, r = w.forest.parent, w.forest.left, w.forest.right
    for _ in range(12):
        u = int(rng.integers(w.forest.n_leaves, w.forest.n_nodes))
        if p[u] >= 0:
            p, l, r = apply_nni(p, l, r, u, int(rng.integers(2)))
    newicks, msa, Q = reference_inputs(Workload(forest_from_parent(p, l, r), w.d, w.qptr, w.qval, w.codes, w.nu, w.phi))
    true_newicks, _, _ = reference_inputs(w)
    t = GPU_EM_solver(true_newicks, {'charMtrx': msa}, {'Q': Q}, {'phi': 0.1, 'nu': 0.1})
    true_score = -t.optimize(initials=1, random_seeds=2, verbose=-1)[0]
    seed(a=3)
    ts = GPU_Topology_search(newicks, GPU_EM_solver, data={'charMtrx': msa}, prior={'Q': Q}, params={'phi': 0.1, 'nu': 0.1})
    ts._open(True)
    start = ts._score_all(deepcopy(DEFAULT_STRATEGY))
    trees, best, params = ts.__search_one__(deepcopy(DEFAULT_STRATEGY), maxiter=25)
    ts.close()
    assert start < true_score - 5                      # the perturbation hurt
    assert best > start + 0.5 * (true_score - start)   # and the search won most of it back


def test_ultrametric_search_goes_candidate_by_candidate():
    """strategy['ultra_constr'] = True (the runner's default): the batch kernel's local EM is unconstrained, so
    candidates are scored one at a time by the reference's own call on the live solver; trees stay ultrametric"""
    from random import seed
    from laml_b200.tree import read_tree_newick
    seed(a=9)
    Q = [dict(q) for q in Q5]
    ts = GPU_Topology_search(['((a,c),(b,d));'], GPU_EM_solver, data={'charMtrx': MSA_A}, prior={'Q': Q}, params={'phi': 0, 'nu': 0})
    st = deepcopy(DEFAULT_STRATEGY)
    st['ultra_constr'] = True
    OracleEngine.instances = 0
    trees, score, params = ts.search(maxiter=4, nreps=1, strategy=st)
    assert OracleEngine.instances == 1 and np.isfinite(score) and ts.trees_scored >= 4
    t = read_tree_newick(trees[0])
    dist = {}
    for nd in t.traverse_preorder():
        dist[nd] = (dist[nd.parent] if nd.parent is not None else 0.0) + (nd.edge_length or 0.0)
    tips = np.array([dist[nd] for nd in t.traverse_leaves()])
    assert np.ptp(tips) < 1e-5 * tips.mean() + 1e-3
    fresh = GPU_EM_solver(trees, {'charMtrx': MSA_A}, {'Q': Q}, params)
    assert abs(-fresh.negative_llh() - score) <= 1e-6 * abs(score)
    # the ultrametric optimum cannot beat the unconstrained brute-force optimum, and the start topology
    # ((a,c),(b,d)) is worse than what the search ends on
    assert -score >= _brute_force(MSA_A, Q) - 1e-6
    s0 = GPU_EM_solver(['((a,c),(b,d));'], {'charMtrx': MSA_A}, {'Q': Q}, {'phi': 0, 'nu': 0})
    assert score >= -s0.optimize(initials=1, random_seeds=4, verbose=-1, ultra_constr=True)[0] - 1e-6


@pytest.mark.parametrize("name", ["ref_rand_small", "ref_rand_forest"])
def test_batch_ultra_scores_equal_the_unclamped_reference_call(name, monkeypatch):
    """``score_neighbourhood(ultra=True)`` = the call the reference makes per candidate (new solver from the
    re-wired newick, ``score_tree`` with ``fixed_brlen`` anchors and ``ultra_constr``) with ONE difference, the
    same as ``clamp_start=False`` on the unconstrained path: lengths are not raised to ``2 dmin`` first.  (With the
    raise, the fixed part of an ultrametric tree that has a branch below ``2 dmin`` stops being ultrametric and the
    reference's local attempt is infeasible by construction -- it then re-optimises every branch,
    Topology_search.py:301-306.)  Scores, and which candidates the constraint leaves no room for, agree."""
    case, z = load_npz_case(name)
    s = _solver(case)
    st = deepcopy(DEFAULT_STRATEGY)
    st['ultra_constr'] = True
    score0, status0 = s.score_tree(strategy=st)                   # an ultrametric starting point
    assert status0 == "optimal"
    from laml_b200.topology import is_ultrametric
    assert is_ultrametric(s._ensure(), s._d())
    nb = score_neighbourhood(s, clamp=False, ultra=True)
    # the clamped call is infeasible whenever a fixed branch sits below 2 dmin
    if (s._d() < 2 * s.dmin - 1e-12).any():
        _, status = score_candidate_like_the_reference(s, int(nb.cand_u[0]), int(nb.cand_k[0]), case, {'ultra_constr': True})
        assert status == "d_infeasible"

    def ini_unclamped(self):
        return [nd.edge_length for tree in self.trees for nd in tree.traverse_postorder()]
    monkeypatch.setattr(GPU_EM_solver, "ini_brlens", ini_unclamped)
    n_nan = 0
    for i in range(len(nb.score)):
        ref_score, status = score_candidate_like_the_reference(s, int(nb.cand_u[i]), int(nb.cand_k[i]), case,
                                                               {'ultra_constr': True})
        if np.isnan(nb.score[i]):
            assert status == "d_infeasible" and nb.iters[i] == -1
            n_nan += 1
        else:
            assert status == "optimal"
            assert abs(ref_score - nb.score[i]) <= 1e-9 * abs(ref_score), (i, ref_score, nb.score[i])
    assert 0 < n_nan < len(nb.score)


def test_batch_ultra_needs_ultrametric_lengths():
    case, z = load_npz_case("ref_rand_small")
    s = _solver(case)
    with pytest.raises(ValueError, match="not ultrametric"):
        score_neighbourhood(s, clamp=False, ultra=True)


@pytest.mark.parametrize("no,T0", [(9, '((a,b),(c,d));'), (10, '(a,b,c,d);')])
def test_reference_search_tests_ultrametric_batched(no, T0):
    """tests 9 and 10 of unit_tests_TopoSearch.py with ``batch_ultra``: neighbourhoods scored in one call under the
    constraint, candidates without room by the reference's sequence; same brute-force optimum to 4 places"""
    from random import seed
    seed(a=1000 + no)
    Q = [dict(q) for q in Q5]
    nllh_bf = _brute_force(MSA_B, Q, ultra_constr=True)
    ts = GPU_Topology_search([T0], GPU_EM_solver, data={'charMtrx': MSA_B}, prior={'Q': Q}, params={'phi': 0, 'nu': 0},
                             batch_ultra=True)
    st = deepcopy(DEFAULT_STRATEGY)
    st['ultra_constr'] = True
    best_trees, max_score, best_params = ts.search(maxiter=200, verbose=False, strategy=st, nreps=1)
    assert abs(nllh_bf + max_score) < 5e-5, (no, -max_score, nllh_bf, best_trees)


@pytest.mark.parametrize("name,rseed,min_sequential", [("ref_rand_small", 77, 0), ("ref_rand_forest", 5, 1)])
def test_batch_ultra_search_on_random_cases(name, rseed, min_sequential):
    """the batched constrained search does not lose score, uses one handle, mixes batch calls with one-at-a-time
    candidates (the forest case meets a candidate without room), and ends on ultrametric, equally high trees whose
    fresh likelihood is the score"""
    from random import seed
    from laml_b200.tree import read_tree_newick
    case, z = load_npz_case(name)
    seed(a=rseed)
    st = deepcopy(DEFAULT_STRATEGY)
    st['ultra_constr'] = True
    ts = GPU_Topology_search(case.newicks, GPU_EM_solver, data={'charMtrx': case.msa}, prior={'Q': case.Q},
                             params={'phi': case.phi, 'nu': case.nu}, batch_ultra=True)
    OracleEngine.instances = 0
    sequential, inner = [], ts._sequential_candidate
    ts._sequential_candidate = lambda *a, **k: (sequential.append(a[:2]), inner(*a, **k))[1]
    trees, score, params = ts.search(maxiter=6, nreps=1, strategy=st)
    assert OracleEngine.instances == 1 and np.isfinite(score) and len(sequential) >= min_sequential
    s0 = _solver(case)
    start, _ = s0.score_tree(strategy=st)
    assert score >= start - 1e-6
    tips = []
    for tr in trees:
        t = read_tree_newick(tr)
        dist = {}
        for nd in t.traverse_preorder():
            dist[nd] = (dist[nd.parent] if nd.parent is not None else 0.0) + (nd.edge_length or 0.0)
        tips += [dist[nd] for nd in t.traverse_leaves()]
    assert np.ptp(np.array(tips)) < 1e-6
    fresh = GPU_EM_solver(trees, {'charMtrx': case.msa}, {'Q': case.Q}, params)
    assert abs(-fresh.negative_llh() - score) <= 1e-6 * abs(score)


def test_full_reoptimisation_per_candidate():
    """local_brlen_opt = False: every parameter re-optimised per candidate on the live handle"""
    from random import seed
    seed(a=10)
    ts = GPU_Topology_search(['((a,c),(b,d));'], GPU_EM_solver, data={'charMtrx': MSA_A}, prior={'Q': [dict(q) for q in Q5]},
                             params={'phi': 0, 'nu': 0})
    st = deepcopy(DEFAULT_STRATEGY)
    st['local_brlen_opt'] = False
    trees, score, params = ts.search(maxiter=20, nreps=1, strategy=st)
    assert abs(_brute_force(MSA_A, [dict(q) for q in Q5]) + score) < 5e-5


@pytest.mark.parametrize("T0,msa", [('((a,b),c,d);', MSA_B), ('(a,b,c,d);', MSA_A), ('((a,b,c),d);', MSA_A)])
def test_search_keeping_polytomies(T0, msa):
    """``search(resolve_polytomies=False)`` on a polytomous input -- the call the runner makes under
    ``--keep_polytomies`` (runner.py:312).  The reference survives it because every candidate gets a fresh solver,
    whose constructor binarises and marks (EM_solver.py:19-43); here ONE solver lives through the search, so it must
    re-resolve the polytomies ``optimize`` collapsed (ML_solver.py:314-333) whenever it re-reads its trees."""
    from random import seed
    seed(a=4242)
    Q = [dict(q) for q in Q5]
    ts = GPU_Topology_search([T0], GPU_EM_solver, data={'charMtrx': msa}, prior={'Q': Q}, params={'phi': 0, 'nu': 0})
    best_trees, max_score, best_params = ts.search(resolve_polytomies=False, maxiter=20, verbose=False,
                                                   strategy=deepcopy(DEFAULT_STRATEGY), nreps=1)
    assert np.isfinite(max_score) and set(best_params) == {'phi', 'nu'}
    # the score is the optimum of the tree that comes back, scored as the reference would: a fresh solver
    s = GPU_EM_solver(best_trees, {'charMtrx': msa}, {'Q': Q}, best_params)
    assert abs(-s.negative_llh() - max_score) <= 1e-6 * abs(max_score)
    # and it is at least as good as the start tree's own optimum (the search only ever accepts or keeps)
    s0 = GPU_EM_solver([T0], {'charMtrx': msa}, {'Q': Q})
    nllh0, _ = s0.optimize(initials=1, verbose=-1)
    assert -max_score <= nllh0 + 1e-4


def test_solver_survives_optimize_on_a_polytomous_tree():
    """negative_llh / score_tree / Estep on the SAME solver after ``optimize`` collapsed its inserted branches"""
    msa = {"a": [1, 0], "b": [1, 2], "c": [0, 2], "d": [1, '?'], "e": [0, 0]}
    Q = [{1: 1.0}, {2: 0.5, 3: 0.5}]
    s = GPU_EM_solver(["((a:0.3,b:0.2,c:0.4)x:0.2,d:0.1,e:0.3)r:0.1;"], {'charMtrx': msa}, {'Q': Q}, {'phi': 0.1, 'nu': 0.1})
    assert s.num_polytomy_mark == 2
    nllh, status = s.optimize(initials=1, random_seeds=3, verbose=-1)
    assert any(len(nd.children) > 2 for nd in s.trees[0].traverse_preorder())        # collapsed, as the reference's
    again = s.negative_llh()                                                            # used to raise: needs binary trees
    assert abs(again - nllh) <= 1e-9 * abs(nllh)
    assert s.num_polytomy_mark == 2 and s.num_edges == 9
    marked = [nd for nd in s.trees[0].traverse_postorder() if nd.polytomy_mark]
    assert len(marked) == 2 and all(abs(nd.edge_length - 5e-5) < 1e-15 for nd in marked)
    score, status = s.score_tree({'ultra_constr': False, 'fixed_phi': None, 'fixed_nu': None, 'fixed_brlen': None})
    assert status == "optimal" and abs(-score - nllh) <= 1e-6 * abs(nllh)
    s.Estep()
