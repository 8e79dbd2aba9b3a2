"""The ultrametric M-step (laml_b200/ultrametric.py) on CPU: KKT / equality checks against the dense
constraint rows of EM_solver.ultrametric_constr (restated in GPU_EM_solver.ultrametric_constr), with
statistics from the oracle."""
import numpy as np
import pytest

from helpers import load_npz_case
from laml_b200 import ultrametric as um
from laml_b200.flatten import flatten_trees
from laml_b200.solver import GPU_EM_solver


def _setup(name, nu=0.1, phi=0.1):
    case, z = load_npz_case(name)
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': phi, 'nu': nu})
    s._forest = flatten_trees(s.trees)                    # no engine: host logic only
    est = case.oracle().estep(case.d, nu, phi)
    S = np.maximum(est["S"].T, 1e-6)
    S = S / S.sum(axis=1, keepdims=True) * s.numsites
    return s, S


def _objective(S, x, nu):
    return float((-(nu + 1) * S[:, 0] * x - nu * (S[:, 1] + S[:, 3]) * x + S[:, 1] * np.log(-np.expm1(-x))
                  + (S[:, 2] + S[:, 4]) * np.log(-np.expm1(-nu * x))).sum())


@pytest.mark.parametrize("name", ["ref_rand_small", "ref_rand_ragged", "ref_rand_forest", "ref_rand_deep"])
def test_solution_is_ultrametric_and_optimal(name):
    s, S = _setup(name)
    prob = s._ultrametric_problem()
    d, H, info = um.solve(prob, S, 0.1)
    assert info["status"] == "optimal"
    M, b = s.ultrametric_constr()                          # the reference's equalities
    fm = s._free_mask().astype(bool)
    assert np.abs(M @ d[fm] - b).max() < 1e-10
    assert (d[fm] >= s.dmin - 1e-12).all() and (d[fm] <= s.dmax + 1e-12).all()
    # first-order optimality on the feasible set: no feasible direction improves the objective.  Feasible
    # directions are height perturbations; check random ones and the coordinate ones.
    rng = np.random.default_rng(0)
    base = _objective(S[fm], d[fm], 0.1)
    n = prob.n
    for _ in range(200):
        dH = np.zeros(n + 1)
        idx = rng.integers(0, n + 1)
        if idx < n and prob.is_leaf[idx]:
            continue
        dH[idx] = rng.choice([-1.0, 1.0]) * 1e-5
        x = (H + dH)[prob.par] - (H + dH)[:n]
        if (x[fm] < s.dmin).any() or (x[fm] > s.dmax).any():
            continue
        assert _objective(S[fm], x[fm], 0.1) <= base + 1e-9


def test_matches_unconstrained_when_inactive():
    """cherries only: a tree with 2 leaves has one equality; compare with a 1-D brute force"""
    s = GPU_EM_solver(["(a:0.3,b:0.7)r:0.2;"], {'charMtrx': {'a': [1, 0, 2], 'b': [1, '?', 0]}},
                      {'Q': [{1: 1.0}, {1: 1.0}, {2: 0.5, 3: 0.5}]}, {'phi': 0.1, 'nu': 0.2})
    s._forest = flatten_trees(s.trees)
    S = np.array([[1.2, 0.9, 0.3, 0.4, 0.2], [0.8, 1.1, 0.2, 0.6, 0.3], [1.5, 0.7, 0.4, 0.2, 0.2]])
    prob = s._ultrametric_problem()
    d, H, info = um.solve(prob, S, 0.2)
    assert abs(d[0] - d[1]) < 1e-12                        # the two leaf branches must be equal
    grid = np.linspace(0.005, 10, 200001)
    vals = [_objective(S[:2], np.array([g, g]), 0.2) for g in grid[::200]]
    g0 = grid[::200][int(np.argmax(vals))]
    fine = np.linspace(max(0.005, g0 - 0.02), g0 + 0.02, 4001)
    best = fine[int(np.argmax([_objective(S[:2], np.array([g, g]), 0.2) for g in fine]))]
    assert abs(d[0] - best) < 2e-5


def test_infeasible_fixed_branches_are_reported():
    s, S = _setup("ref_rand_small")
    nodes = s._forest.nodes
    leaves = [nd for nd in nodes if nd.is_leaf()]
    # two sibling leaves with different fixed lengths cannot be ultrametric
    sib = next(nd for nd in nodes if not nd.is_leaf() and all(c.is_leaf() for c in nd.children))
    a, b = sib.children
    a.mark_fixed = b.mark_fixed = True
    a.edge_length, b.edge_length = 0.1, 0.4
    prob = s._ultrametric_problem()
    d, H, info = um.solve(prob, S, 0.1)
    assert d is None and info["status"] == "infeasible"


@pytest.mark.parametrize("name", ["ref_rand_small", "ref_rand_ragged", "ref_rand_forest", "ref_rand_deep", "ref_example1_start"])
def test_native_solver_matches_the_numpy_model(name):
    """pmm_ultrametric_solve (C++, laml_b200/csrc/pmm_ultrametric.cpp) against the numpy statement of the same
    algorithm: cold start, warm start from the previous heights with perturbed statistics, nu below eps_nu,
    fixed and polytomy-marked (zero) branches"""
    s, S = _setup(name)
    rng = np.random.default_rng(5)
    nodes = s._forest.nodes
    internal = [nd for nd in nodes if nd.children and nd.parent is not None]
    for variant in ("plain", "fixed", "zero"):
        for nd in nodes:
            nd.mark_fixed, nd.polytomy_mark = False, False
        if variant == "fixed":
            # fix a few internal branches at their current lengths (what fixed_brlen does, ML_solver.py:296-305)
            for nd in internal[::4]:
                nd.mark_fixed = True
                nd.edge_length = float(max(nd.edge_length, 0.02))
        if variant == "zero" and internal:
            internal[len(internal) // 2].polytomy_mark = True
        prob = s._ultrametric_problem()
        for nu in (0.1, 0.0):
            d0, H0, i0 = um.solve_numpy(prob, S, nu)
            d1, H1, i1 = um.solve(prob, S, nu)
            if d0 is None:
                assert d1 is None and i1["status"] == "infeasible" == i0["status"]
                continue
            assert i0["status"] == i1["status"] == "optimal"
            # both stop inside the same solver tolerance (barrier weight 1e-11, Newton decrement test): ~1e-8
            np.testing.assert_allclose(d1, d0, rtol=0, atol=5e-8)
            np.testing.assert_allclose(H1, H0, rtol=0, atol=5e-8)
            S2 = S * rng.uniform(0.9, 1.1, S.shape)
            dw0, Hw0, iw0 = um.solve_numpy(prob, S2, nu, H0=H0)
            dw1, Hw1, iw1 = um.solve(prob, S2, nu, H0=H1)
            np.testing.assert_allclose(dw1, dw0, rtol=0, atol=5e-8)
            dc, _, ic = um.solve(prob, S2, nu)                       # cold start: same optimum, more steps
            np.testing.assert_allclose(dw1, dc, rtol=0, atol=5e-8)
            assert iw1["newton_iters"] < ic["newton_iters"]


def test_native_solver_usage_errors():
    from laml_b200.capi import PMMError, ultrametric_solve
    parent = np.array([2, 2, -1], np.int32)
    with pytest.raises(PMMError):
        ultrametric_solve(parent, np.zeros(3, np.int8), np.zeros(3), 0.005, 10.0, np.ones((2, 5)), 0.1)
    with pytest.raises(PMMError):
        ultrametric_solve(np.array([1, 0, -1], np.int32), np.zeros(3, np.int8), np.zeros(3), 0.005, 10.0, np.ones((3, 5)), 0.1)   # a cycle
    d, H, rc, it = ultrametric_solve(parent, np.zeros(3, np.int8), np.zeros(3), 0.005, 10.0, np.ones((3, 5)), 0.1)
    assert rc == 0 and abs(d[0] - d[1]) < 1e-12 and H[0] == H[1] == 0.0
