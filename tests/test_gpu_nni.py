"""Batched NNI candidate scoring against the engine's own full-tree EM on the
explicitly re-wired topology with five free branches -- what the reference's apply_nni + score_tree
compute (Topology_search.py:242-322, ML_solver.py:114-123)."""
import numpy as np
import pytest

from laml_b200.capi import Engine
from laml_b200.nni import apply_nni, nni_candidates
from laml_b200.synth import make_workload

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["records", "rows"])
def nni_cache(request, monkeypatch):
    """the batch scorer on the lumped records (default) and on cached full-state rows (PMM_NNI_ROWS=1)"""
    monkeypatch.setenv("PMM_NNI_ROWS", "1" if request.param == "rows" else "0")
    return request.param


@pytest.mark.parametrize("seed,n,K,amax,nu,phi", [(71, 40, 14, 12, 0.15, 0.1), (72, 25, 9, 40, 0.0, 0.2), (73, 60, 20, 6, 0.3, 0.0)])
def test_batch_matches_full_tree_em(seed, n, K, amax, nu, phi):
    w = make_workload(n, K, 1, amax, max(nu, 0.1), max(phi, 0.1), seed=seed)
    codes = w.codes.copy()
    if phi == 0.0:
        pass
    d = np.maximum(w.d, 0.01)                       # ini_brlens' clamp (ML_solver.py:195-203)
    f = w.forest
    eng = Engine(n, w.qptr, w.qval, codes)
    eng.set_tree(f.parent, f.left, f.right)
    cu, ck = nni_candidates(f.parent, f.left, n)
    res = eng.score_nni_batch(cu, ck, d, nu, phi)
    assert len(cu) == 2 * (n - 2)
    R = float((codes != -1).sum())
    ref_eng = Engine(n, w.qptr, w.qval, codes)
    checked_root = False
    for i in range(len(cu)):
        p2, l2, r2 = apply_nni(f.parent, f.left, f.right, int(cu[i]), int(ck[i]))
        ref_eng.set_tree(p2, l2, r2)
        free = np.zeros(f.n_nodes, np.uint8)
        free[res["nodes"][i]] = 1
        ref = ref_eng.em(d, nu, phi, R, n, free_mask=free, optimize_nu=False, optimize_phi=False)
        assert np.isfinite(ref["nllh"])
        assert abs(res["nllh"][i] - ref["nllh"]) <= 1e-9 * abs(ref["nllh"]), (i, res["nllh"][i], ref["nllh"])
        assert res["iters"][i] == ref["iters"]
        np.testing.assert_allclose(res["d5"][i], ref["d"][res["nodes"][i]], rtol=1e-6, atol=1e-8)
        untouched = np.ones(f.n_nodes, bool); untouched[res["nodes"][i]] = False
        assert (ref["d"][untouched] == d[untouched]).all()
        checked_root |= f.parent[f.parent[cu[i]]] < 0
    assert checked_root                              # moves right under the root were covered
    ref_eng.close(); eng.close()


def test_batch_on_forest_and_repeatability():
    from test_gpu_edge_cases import test_forest_of_two_trees  # noqa: F401  (layout helper lives there)
    w = make_workload(30, 10, 2, 10, 0.2, 0.2, seed=81)
    d = np.maximum(w.d, 0.01)
    eng = Engine(30, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    cu, ck = nni_candidates(w.forest.parent, w.forest.left, 30)
    a = eng.score_nni_batch(cu, ck, d, 0.2, 0.2)
    b = eng.score_nni_batch(cu[::-1].copy(), ck[::-1].copy(), d, 0.2, 0.2)
    assert (a["nllh"] == b["nllh"][::-1]).all() and (a["d5"] == b["d5"][::-1]).all()      # bit-reproducible, order-free
    base = -eng.inside(d, 0.2, 0.2)
    assert (a["nllh"] < base + 50).all()
    eng.close()
