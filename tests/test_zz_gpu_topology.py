"""NNI search on the batched scorer through the real engine: the batch entry of every
candidate equals the call the reference's ``apply_nni`` makes for it (new solver from the re-wired newick,
``score_tree`` with ``fixed_brlen`` anchors -- tests/topo_reference_call.py), and the reference's own search
tests (laml_unit_tests/unit_tests_TopoSearch.py) reach their brute-force optima.  Same tests as
tests/test_topology_host.py, which runs them on CPU against the oracle."""
from copy import deepcopy

import numpy as np
import pytest

from helpers import load_npz_case
from laml_b200.solver import DEFAULT_STRATEGY, GPU_EM_solver
from laml_b200.topology import GPU_Topology_search, apply_move, score_neighbourhood
from test_topology_host import MSA_A, MSA_B, Q5, _brute_force
from topo_reference_call import score_candidate_like_the_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ref_rand_small", "ref_rand_forest"])
def test_batch_scores_equal_the_reference_call(name):
    case, z = load_npz_case(name)
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': case.phi, 'nu': case.nu})
    nb = score_neighbourhood(s)
    assert s._engine.info()["kernel_launches"] > 0
    for i in range(len(nb.score)):
        ref_score, status = score_candidate_like_the_reference(s, int(nb.cand_u[i]), int(nb.cand_k[i]), case)
        assert status == "optimal"
        assert abs(ref_score - nb.score[i]) <= 1e-8 * abs(ref_score), (i, ref_score, nb.score[i])
    i = int(np.argmax(nb.score))
    eng = s._engine
    apply_move(s, nb, i)
    assert s._engine is eng
    assert abs(-s.negative_llh() - nb.score[i]) <= 1e-8 * abs(nb.score[i])
    s.close()


@pytest.mark.parametrize("no,msa,T0,resolve_only,nllh_bf", [
    (1, MSA_A, '((a,c),(b,d));', False, None),
    (5, MSA_A, '((a,c),(b,d));', True, 11.809140958825493),
    (6, MSA_B, '((a,b),(c,d));', False, None),
    (7, MSA_B, '((a,b),c,d);', True, 10.083048489886435),
    (8, MSA_B, '((c,d),a,b);', True, 10.08306053772235),
])
def test_reference_search_tests(no, msa, T0, resolve_only, nllh_bf):
    from random import seed
    seed(a=1000 + no)
    Q = [dict(q) for q in Q5]
    if nllh_bf is None:
        nllh_bf = _brute_force(msa, Q)
    ts = GPU_Topology_search([T0], GPU_EM_solver, data={'charMtrx': msa}, prior={'Q': Q}, params={'phi': 0, 'nu': 0})
    st = deepcopy(DEFAULT_STRATEGY)
    st['resolve_search_only'] = resolve_only
    best_trees, max_score, best_params = ts.search(maxiter=200, verbose=False, strategy=st, nreps=1)
    assert abs(nllh_bf + max_score) < 5e-5, (no, -max_score, nllh_bf, best_trees)
