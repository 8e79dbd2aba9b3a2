"""Posterior annotation through the real engine: ``laml_b200.annotate.annotate`` against
the files the reference's own runner code wrote (tests/golden/make_annotation_golden.py).  The engine's
posteriors differ from the reference's in the 12th digit, so a printed 2-digit probability may round the
other way at a tie: structure, labels and listed states must agree, numbers within one printed unit
(tests/annot_compare.py)."""
import os
import sys

import pytest

from annot_compare import assert_annotations_match
from helpers import GOLDEN, load_npz_case
from laml_b200.annotate import annotate

sys.path.insert(0, GOLDEN)
from make_annotation_golden import CASES, POLYTOMY, strip_internal_labels  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_annotation_file_matches_the_reference(name, tmp_path):
    npz, strip = CASES[name]
    case, z = load_npz_case(npz)
    trees = [strip_internal_labels(t) if strip else t for t in case.newicks]
    out = tmp_path / "annot.txt"
    annotate(trees, case.msa, case.Q, {"phi": case.phi, "nu": case.nu}, str(out))
    assert_annotations_match(out.read_text(), open(os.path.join(GOLDEN, "annot_%s.txt" % name)).read(), exact=False)


def test_annotation_polytomy_and_raw_prior(tmp_path):
    out = tmp_path / "annot.txt"
    annotate(POLYTOMY["trees"], POLYTOMY["msa"], POLYTOMY["Q"], {"phi": POLYTOMY["phi"], "nu": POLYTOMY["nu"]}, str(out))
    assert_annotations_match(out.read_text(), open(os.path.join(GOLDEN, "annot_polytomy.txt")).read(), exact=False)


def test_published_example2_annotation_file(tmp_path):
    """the annotation file of the reference's published example2 run (examples/out_example2, LAML 1.0.0 with MOSEK
    and treeswift), reproduced byte for byte on CPU (tests/test_published_runs.py), through the engine"""
    from test_published_runs import annotation_case
    trees, msa, Q, pub = annotation_case()
    out = tmp_path / "annot.txt"
    annotate(trees, msa, Q, {"phi": pub["phi"], "nu": pub["nu"]}, str(out))
    assert_annotations_match(out.read_text(), open(os.path.join(GOLDEN, "annot_example2_published.txt")).read(), exact=False)


def test_published_example1_annotation_file(tmp_path):
    """examples/out_example1/example1_annotations.txt through the engine (byte for byte on CPU, tests/test_published_runs.py)"""
    case, z = load_npz_case("ref_example1_optimum")
    trees = [t.replace("[&R] ", "") for t in case.newicks]
    out = tmp_path / "annot.txt"
    annotate(trees, case.msa, case.Q, {"phi": case.phi, "nu": case.nu}, str(out))
    assert_annotations_match(out.read_text(), open(os.path.join(GOLDEN, "annot_example1_published.txt")).read(), exact=False)
