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
