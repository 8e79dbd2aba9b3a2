"""Dataflow launches of the lumped engine on the GPU (PMM_DF=1 / 2: several tiers of the clade schedule in ONE launch,
CTAs ordered by atomic tickets and synchronised through per-(clade, site block) marks, laml_b200/csrc/pmm_lumped.cuh)
against the launch-per-tier schedule: the same bits, E-step after E-step (marks and tickets carry over), on one and
several site blocks, a forest, and through the EM loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _estep_bits(w, env, monkeypatch, reps=3, caps=None):
    from laml_b200.capi import Engine
    for k in ("PMM_DF", "PMM_CLADE", "PMM_CLADE_TOP"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    if caps:
        monkeypatch.setenv("PMM_CLADE", str(caps[0])); monkeypatch.setenv("PMM_CLADE_TOP", str(caps[1]))
    eng = Engine(w.forest.n_leaves, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    outs = []
    d = w.d
    for r in range(reps):
        outs.append(eng.estep(d * (1.0 + 0.1 * r), w.nu, w.phi, want_site_llh=True))
    info = eng.info()
    R = float((np.asarray(w.codes) != -1).sum())
    em = eng.em(np.maximum(w.d, 0.01), w.nu, w.phi, R, w.forest.n_leaves, max_iter=15)
    eng.close()
    return outs, info, em


@pytest.mark.parametrize("n,K,caps", [(500, 40, (8, 4)), (900, 700, (16, 6)), (3000, 300, (40, 3)), (64, 129, (2, 2))])
def test_dataflow_launches_give_the_same_bits(monkeypatch, n, K, caps):
    from laml_b200.synth import make_workload
    w = make_workload(n, K, 1, 25, 0.2, 0.25, seed=n + K)
    base, info0, em0 = _estep_bits(w, {"PMM_DF": "0"}, monkeypatch, caps=caps)
    assert info0["launches_per_pass"] >= 3
    for mode, launches in (("1", 2), ("2", 1)):
        got, info, em = _estep_bits(w, {"PMM_DF": mode}, monkeypatch, caps=caps)
        assert info["launches_per_pass"] == launches and info["clades"] == info0["clades"]
        for a, b in zip(got, base):
            assert np.isfinite(a["llh"]) and a["llh"] == b["llh"] and a["Rtilde"] == b["Rtilde"]
            assert a["S"].tobytes() == b["S"].tobytes() and a["site_llh"].tobytes() == b["site_llh"].tobytes()
        assert em["iters"] == em0["iters"] and em["nllh"] == em0["nllh"] and em["d"].tobytes() == em0["d"].tobytes()
