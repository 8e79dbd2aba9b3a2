"""Posterior annotation on CPU: the array-level tag functions against the reference's
``az_partition`` dumps, and the writer against files produced by the reference's own runner code
(tests/golden/make_annotation_golden.py), with the engine swapped for the oracle-backed stand-in
(tests/oracle_engine.py -- test infrastructure only).  tests/test_gpu_annotate.py runs the same files
through the real engine."""
import os

import numpy as np
import pytest

import laml_b200.solver as solver_mod
from annot_compare import assert_annotations_match
from helpers import GOLDEN, load_npz_case
from laml_b200.annotate import TAG_MISSING, TAG_Z, annotate, az_tags, fill_missing_tags, format_posterior
from oracle_engine import OracleEngine

import sys
sys.path.insert(0, GOLDEN)
from make_annotation_golden import CASES, POLYTOMY, strip_internal_labels  # noqa: E402


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)


@pytest.mark.parametrize("name", ["ref_rand_small", "ref_rand_ragged", "ref_rand_forest", "ref_rand_deep",
                                  "ref_test4_interior", "ref_rand_zeroprior"])
def test_az_tags_match_the_reference_partition(name):
    """against node.alpha dumped from the reference's own az_partition (tests/golden/make_golden.py:
    'z' -> -1, '?' -> -2, otherwise the original state id)"""
    case, z = load_npz_case(name)
    ids = case.ids(z["labels"], z["tree_of"])
    got = az_tags(case.forest, case.enc.codes[case.leaf_row])[ids]
    enc = case.enc
    as_state = np.empty_like(z["tag"])
    for j in range(enc.n_sites):
        col = got[:, j]
        lut = np.concatenate([[-2, -1], enc.states[enc.qptr[j]:enc.qptr[j + 1]]])       # TAG_MISSING, TAG_Z, ranks 1..
        as_state[:, j] = lut[col + 1]
    assert (as_state == z["tag"]).all()


def test_fill_missing_tags_small():
    from laml_b200.flatten import flatten_trees
    from laml_b200.tree import read_tree_newick
    f = flatten_trees([read_tree_newick("((a:1,b:1)e:1,(c:1,d:1)f:1)r:1;")])
    codes = np.array([[2, -1, 0], [-1, -1, 1], [-1, -2, 1], [-1, -1, -1]], np.int32)       # a, b, c, d
    tags = az_tags(f, codes)
    lab = {nd.label: i for i, nd in enumerate(f.nodes)}
    assert tags[lab["e"]].tolist() == [2, TAG_MISSING, TAG_Z]
    assert tags[lab["f"]].tolist() == [TAG_MISSING, TAG_MISSING, 1]
    assert tags[lab["r"]].tolist() == [2, TAG_MISSING, TAG_Z]
    filled = fill_missing_tags(f, tags)
    assert filled[lab["f"]].tolist() == [2, TAG_MISSING, 1]          # '?' below alpha takes alpha
    assert filled[lab["d"]].tolist() == [2, TAG_MISSING, 1]          # ... all the way down
    assert filled[lab["b"]].tolist() == [2, TAG_MISSING, 1]          # below 'z' it keeps its own tag


def test_format_posterior_forms():
    q = {0: 0.0, 1: 2.0, 2: 1.0}
    assert format_posterior(1.0, 0.0, 0.0, '5', q) == '0'
    assert format_posterior(0.0, 1.0, 0.0, '5', q) == '-1'
    assert format_posterior(0.0, 0.0, 1.0, '5', q) == '5'
    assert format_posterior(0.3, 0.51, 0.19, '?', q) == '0:0.3/-1:0.51/1:0.38/2:0.19'     # raw prior, as the runner
    assert format_posterior(0.25, 0.0, 0.75, '7', q) == '0:0.25/7:0.75'
    assert format_posterior(0.99, 0.0, 0.01, '?', {3: 0.01}) == '0:0.99/'                 # the reference's quirk


@pytest.mark.parametrize("name", sorted(CASES))
def test_writer_reproduces_the_reference_file(name, tmp_path):
    npz, strip = CASES[name]
    case, z = load_npz_case(npz)
    trees = [strip_internal_labels(t) if strip else t for t in case.newicks]
    out = tmp_path / "annot.txt"
    annotate(trees, case.msa, case.Q, {"phi": case.phi, "nu": case.nu}, str(out))
    assert_annotations_match(out.read_text(), open(os.path.join(GOLDEN, "annot_%s.txt" % name)).read(), exact=True)


def test_writer_polytomy_and_raw_prior(tmp_path):
    out = tmp_path / "annot.txt"
    annotate(POLYTOMY["trees"], POLYTOMY["msa"], POLYTOMY["Q"], {"phi": POLYTOMY["phi"], "nu": POLYTOMY["nu"]}, str(out))
    assert_annotations_match(out.read_text(), open(os.path.join(GOLDEN, "annot_polytomy.txt")).read(), exact=True)


def test_writer_leaves_the_solver_untouched(tmp_path):
    from laml_b200.annotate import write_annotations
    from laml_b200.solver import GPU_EM_solver
    case, z = load_npz_case("ref_rand_forest")
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': case.phi, 'nu': case.nu})
    before, nllh = s.get_tree_newick(), s.negative_llh()
    write_annotations(s, case.Q, str(tmp_path / "a.txt"))
    assert s.get_tree_newick() == before and s.negative_llh() == nllh
